#!/usr/bin/env python
"""Reflow sampling (`rectify_stage.dosampling`) at the reference's size: 65 536 samples x 64 wi =
4 194 304 rows, forward-only Euler, T = 256 (disk, 32-wide net) / 128 (spherical, 64-wide 6-hidden teacher net).
Times bsdfdiff_flow_forward on the tensor-core path and on the fp32 CUDA-core path (CUDA events, best of 3)."""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                    # noqa: E402
import bsdf_diffusion_sampling_b200 as pkg      # noqa: E402
from oracle import bsdf_oracle as O             # noqa: E402

G = os.path.join(ROOT, "tests", "golden")
for name, fn, T, macs in (("disk_w32", "disk_aniso_brushed_aluminium_1_rgb.npz", 256, 2912),
                          ("spherical_w64", "spherical_aniso_brushed_aluminium_1_rgb.npz", 128, 22272)):
    flow, base, z = O.load_material_npz(os.path.join(G, fn))
    if "reflow_w0" in z:
        flow = O.FlowWeights([z[f"reflow_w{i}"] for i in range(int(z["n_reflow_layers"]))])
    pf = pkg.weights.pack_flow_layers(flow.layers, "cuda")
    rng = np.random.default_rng(0)
    n = 65536 * 64
    wi = torch.from_numpy(np.repeat(z["reflow_wi"][:64], 65536, 0).astype(np.float32)).cuda()
    x0 = torch.from_numpy(rng.normal(0, 0.3, (n, 2)).astype(np.float32)).cuda()
    for prec in ("tc16", "fp32"):
        reps = 3 if prec == "tc16" else 1
        pkg.ops.flow_forward(wi[:1 << 16], pf, T, x0=x0[:1 << 16], precision=prec)
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            x, _ = pkg.ops.flow_forward(wi, pf, T, x0=x0, precision=prec)
            e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        flops = 2.0 * macs * T * n
        print(json.dumps({"net": name, "precision": prec, "rows": n, "T": T, "ms": best, "rows_per_s": n / best * 1e3,
                          "algorithmic_tflops": flops / best / 1e9,
                          "finite": bool(torch.isfinite(x).all().item())}), flush=True)
