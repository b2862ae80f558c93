#!/usr/bin/env python
"""Error statistics of a precision path against the committed goldens (reference outputs) -- the numbers
behind the tolerance table in DESIGN.md -- for several thresholds of the conditioning-triggered fp32 fix-up
(0 = the tensor-core kernel alone).  RAW relative errors; the conditioning-weighted figure is a diagnostic.  Run on the GPU box:
    python profiles/tc_error_stats.py [tc16|tc16_exp|fp32 ...] > gpurun_out/err.txt"""
import glob
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bsdf_diffusion_sampling_b200 as pkg  # noqa: E402
from oracle import bsdf_oracle as O  # noqa: E402
from oracle import c_oracle as C  # noqa: E402


def q(a):
    a = a[np.isfinite(a)]
    return f"med {np.median(a):.2e} p99 {np.quantile(a, 0.99):.2e} max {a.max():.2e}"


def main():
    precs = sys.argv[1:] or ["tc16"]
    for path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz"))):
        flow, base, z = O.load_material_npz(path)
        pf = pkg.weights.pack_flow_layers(flow.layers, "cuda")
        pb = pkg.weights.pack_base_arrays(base.w1, base.b1, base.wo, base.bo, "cuda")
        T = int(z["T"])
        wi, x0 = torch.from_numpy(z["wi"]).cuda(), torch.from_numpy(z["x0"]).cuda()
        _, _, md_s = C.sample(flow, base, z["wi"], T, z["x0"], with_mindet=True)
        _, md_p = C.pdf(flow, base, z["wo_eval"], z["wi_eval"], T, with_mindet=True)
        n = z["wi"].shape[0]
        for prec in precs:
            for thr in ((0.0,) if prec == "fp32" else (0.0, 0.1, 0.25, 0.5)):
                x, pdf, _ = pkg.ops.sample(wi, pf, pb, T, x0=x0, precision=prec, fixup=thr)
                ns = pkg.ops.last_fixup_count() if thr else 0
                p = pkg.ops.pdf(torch.from_numpy(z["wo_eval"]).cuda(), torch.from_numpy(z["wi_eval"]).cuda(), pf, pb, T,
                                precision=prec, fixup=thr).cpu().numpy()
                np_ = pkg.ops.last_fixup_count() if thr else 0
                x, pdf = x.cpu().numpy(), pdf.cpu().numpy()
                rs = np.abs(pdf - z["pdf_sample"]) / np.maximum(np.abs(z["pdf_sample"]), 1e-6)
                rp = np.abs(p - z["pdf_eval"]) / np.maximum(np.abs(z["pdf_eval"]), 1e-6)
                print(f"{os.path.basename(path)[:-4]:44s} {prec:8s} T={T} fixup<{thr}: recomputed in fp32 "
                      f"{100.0 * ns / n:.2f}% (sample) {100.0 * np_ / z['wo_eval'].shape[0]:.2f}% (pdf)")
                print(f"    |dx|                 {q(np.abs(x - z['x']).ravel())}")
                print(f"    sample pdf rel       {q(rs)}   weighted {q(rs * md_s)}")
                print(f"    pdf()  rel           {q(rp)}   weighted {q(rp * md_p)}")
    print("timeout flag:", pkg._lib.lib.bsdfdiff_debug_timeout_flag())


if __name__ == "__main__":
    main()
