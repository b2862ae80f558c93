#!/usr/bin/env python
"""Diagnostic: run-to-run determinism of the tensor-core pdf()/sample() kernels and agreement between the shipped TMEM map and
the -DBSDFDIFF_TC_NOALIAS build (GPU box).  usage: python profiles/diag_noalias.py"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bsdf_diffusion_sampling_b200 as pkg  # noqa: E402
from oracle import bsdf_oracle as O  # noqa: E402

L = pkg._lib
alt = ctypes.CDLL(os.path.join(ROOT, "variants", "lib_noalias.so"))
alt.bsdfdiff_sample.argtypes = L.lib.bsdfdiff_sample.argtypes
alt.bsdfdiff_pdf.argtypes = L.lib.bsdfdiff_pdf.argtypes


def run_pdf(lib, pf, pb, T, wo, wi):
    out = torch.empty(wi.shape[0], device="cuda")
    rc = lib.bsdfdiff_pdf(L.PREC_TC16, pf.domain, 0, T, wi.shape[0], wo.data_ptr(), wi.data_ptr(), pf.blob.data_ptr(), pf.hidden,
                          pf.n_hidden, pb.data_ptr(), out.data_ptr(), 0.0, None, torch.cuda.current_stream().cuda_stream)
    assert rc == 0
    torch.cuda.synchronize()
    return out


for name in ("disk_aniso_brushed_aluminium_1_rgb", "bsdf_0", "spherical_ilm_solo_m_68_rgb"):
    flow, base, z = O.load_material_npz(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    pf = pkg.weights.pack_flow_layers(flow.layers, "cuda")
    pb = pkg.weights.pack_base_arrays(base.w1, base.b1, base.wo, base.bo, "cuda")
    T = int(z["T"])
    for reps in (1, 64):
        wo = torch.from_numpy(np.tile(z["wo_eval"], (reps, 1))).cuda()
        wi = torch.from_numpy(np.tile(z["wi_eval"], (reps, 1))).cuda()
        a = [run_pdf(L.lib, pf, pb, T, wo, wi) for _ in range(3)]
        b = [run_pdf(alt, pf, pb, T, wo, wi) for _ in range(3)]
        f32 = pkg.ops.pdf(wo, wi, pf, pb, T, precision="fp32")
        print(f"{name} n={wi.shape[0]}: shipped run-to-run diffs {[int((a[0] != x).sum()) for x in a[1:]]}, "
              f"noalias run-to-run {[int((b[0] != x).sum()) for x in b[1:]]}, shipped vs noalias {int((a[0] != b[0]).sum())}")
        d = (a[0] != b[0]).nonzero().squeeze(1)
        if d.numel():
            i = d[:12].cpu().numpy()
            ra = ((a[0] - f32).abs() / f32.abs().clamp_min(1e-6))[d]
            rb = ((b[0] - f32).abs() / f32.abs().clamp_min(1e-6))[d]
            print("   first differing rows", i.tolist(), "row%128", (i % 128).tolist(), "tile", (i // 128).tolist())
            print("   shipped", a[0][d[:6]].tolist(), "\n   noalias", b[0][d[:6]].tolist(), "\n   fp32   ", f32[d[:6]].tolist())
            print(f"   rel err vs fp32 on the differing rows: shipped median {ra.median():.2e} max {ra.max():.2e}; noalias median {rb.median():.2e} max {rb.max():.2e}")
