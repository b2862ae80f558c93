#!/usr/bin/env python
"""Fixtures for the measured-BSDF evaluator (run in the build container, where /root/reference exists):

    python tests/golden/make_measured_golden.py

For two isotropic RGL materials that also have network goldens (vch_silk_blue_rgb: disk plugin, chm_mint_rgb: spherical plugin)
it stores the raw tensor-file fields the model needs (theta_i, phi_i, ndf, sigma, vndf, rgb, jacobian -- measurement DATA of the
RGL material database, rendering/measuredbsdfs/<mat>.bsdf, 0.66 MB each) plus 4096 random (wi, wo) pairs and the value the numpy
oracle (oracle/measured_oracle.py) gives for them.  The oracle restates Mitsuba 3's `measured` eval, which cannot run here
(parity with Mitsuba itself is unpinned; see the oracle's header for what pins it instead).  The anisotropic materials (4.5 MB
each, 17 phi_i slices) are not stored: the parameter-axis interpolation they exercise is covered by a synthetic anisotropic
tensor file the tests build, and by tests that read /root/reference when it exists."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import measured_oracle as M  # noqa: E402

SRC = "/root/reference/rendering/measuredbsdfs"
KEEP = ("theta_i", "phi_i", "ndf", "sigma", "vndf", "rgb", "jacobian")


def main():
    for mat in ("vch_silk_blue_rgb", "chm_mint_rgb"):
        f = M.read_tensor_file(os.path.join(SRC, mat + ".bsdf"))
        b = M.MeasuredBSDF(f)
        rng = np.random.default_rng(7)
        w = rng.normal(size=(2, 4096, 3))
        w[:, :, 2] = np.abs(w[:, :, 2]) + 0.02
        w /= np.linalg.norm(w, axis=2, keepdims=True)
        w[1, ::11, 2] *= -1.0                                     # some wo below the horizon: eval must return 0
        wi, wo = w[0].astype(np.float32), w[1].astype(np.float32)
        val = b.eval(wi, wo).astype(np.float32)
        out = {k: np.asarray(f[k]) for k in KEEP}
        out.update(wi=wi, wo=wo, eval=val, material=mat)
        path = os.path.join(ROOT, "tests", "golden", "measured", f"measured_{mat}.npz")
        np.savez_compressed(path, **out)
        print(path, os.path.getsize(path) // 1024, "KiB; eval range", float(val.min()), float(val.max()))


if __name__ == "__main__":
    main()
