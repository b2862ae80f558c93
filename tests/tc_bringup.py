"""Bring-up helper (not a pytest): run the tcgen05 path on the goldens under the layout debug knobs and
print error statistics.  Usage on the GPU box:  timeout 300 python tests/tc_bringup.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bsdf_diffusion_sampling_b200 as pkg  # noqa: E402
from oracle import bsdf_oracle as O  # noqa: E402


def stats(name, a, b):
    err = np.abs(a - b)
    print(f"   {name}: median {np.median(err):.3e} p99 {np.quantile(err, 0.99):.3e} max {np.nanmax(err):.3e} "
          f"nan {np.isnan(a).sum()}")


def main():
    knobs = [int(k) for k in (sys.argv[1:] or ["0", "1", "2", "3"])]
    for fname in ("disk_aniso_brushed_aluminium_1_rgb", "spherical_aniso_brushed_aluminium_1_rgb"):
        flow, base, z = O.load_material_npz(os.path.join(ROOT, "tests", "golden", fname + ".npz"))
        pf = pkg.weights.pack_flow_layers(flow.layers, "cuda")
        pb = pkg.weights.pack_base_arrays(base.w1, base.b1, base.wo, base.bo, "cuda")
        wi, x0 = torch.from_numpy(z["wi"]).cuda(), torch.from_numpy(z["x0"]).cuda()
        T = int(z["T"])
        for dbg in knobs:
            os.environ["BSDFDIFF_TC_DEBUG"] = str(dbg)
            for prec in ("tc16_exp", "tc16"):
                x, pdf, _ = pkg.ops.sample(wi, pf, pb, T, x0=x0, precision=prec)
                torch.cuda.synchronize()
                print(f"{fname} debug={dbg} {prec}: timeout_flag={pkg._lib.lib.bsdfdiff_debug_timeout_flag()}")
                stats("x  ", x.cpu().numpy(), z["x"])
                p = pdf.cpu().numpy()
                rel = np.abs(p - z["pdf_sample"]) / np.maximum(np.abs(z["pdf_sample"]), 1e-6)
                print(f"   pdf rel: median {np.nanmedian(rel):.3e} p99 {np.nanquantile(rel, 0.99):.3e}")
            # T = 1 isolates a single network evaluation
            x1, _, _ = pkg.ops.sample(wi, pf, pb, 1, x0=x0, precision="tc16_exp")
            xr, _, _ = pkg.ops.sample(wi, pf, pb, 1, x0=x0, precision="fp32")
            stats("T=1 x vs fp32 kernel", x1.cpu().numpy(), xr.cpu().numpy())
        os.environ["BSDFDIFF_TC_DEBUG"] = "0"


if __name__ == "__main__":
    main()
