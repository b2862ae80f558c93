#!/usr/bin/env python
"""In-kernel phase timers of the tcgen05 flow kernel (BSDFDIFF_TC_PROFILE=1 instantiation).

Prints, per warp role, the share of cycles spent in each phase.  Run on the GPU box:
    BSDFDIFF_TC_PROFILE=1 python profiles/tc_phase_profile.py [disk|spherical] > gpurun_out/phase.txt
"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("BSDFDIFF_TC_PROFILE", "1")
import bench  # noqa: E402
import bsdf_diffusion_sampling_b200 as pkg  # noqa: E402

workload = sys.argv[1] if len(sys.argv) > 1 else "disk"
T = 4 if workload == "disk" else 8
layers, base = bench.load_fixture(workload)
pf = pkg.weights.pack_flow_layers(layers, "cuda")
pb = pkg.weights.pack_base_arrays(*base, "cuda")
s = pkg.plugins.NeuralBSDFSampler(workload, pf, pb, T=T, precision="tc16")
wi = torch.from_numpy(bench.synth_wi3(workload, 4096, 1)).cuda()
for _ in range(2):
    s.sample(wi, seed=1)
torch.cuda.synchronize()
n = 148 * 12 * 8
buf = (ctypes.c_ulonglong * n)()
got = pkg._lib.lib.bsdfdiff_debug_profile_fetch(buf, n)
assert got == n, got
a = np.frombuffer(buf, dtype=np.uint64).reshape(148, 12, 8).astype(np.float64)
names = ["prologue+state", "wait MMA", "ld+math+st", "wait::st+barrier", "MMA issue", "output+epilogue", "total", "-"]
print(f"workload {workload} T={T} 16.7M queries; cycles per warp (mean over 148 CTAs)")
tot = a[:, :, 6].mean()
print(f"total cycles/warp {tot:.0f}")
for role, sel in (("issuer warps (q=0)", [0, 4, 8]), ("other warps", [1, 2, 3, 5, 6, 7, 9, 10, 11])):
    m = a[:, sel, :].mean(axis=(0, 1))
    print(role)
    for k in range(6):
        print(f"   {names[k]:18s} {m[k]:12.0f}  {100 * m[k] / m[6]:5.1f} %")
    print(f"   accounted {100 * m[:6].sum() / m[6]:.1f} %")
tiles_per_group = (4096 * 4096 / 128) / 148 / 3
rounds = tiles_per_group * T * (len(layers))
print(f"tiles/group {tiles_per_group:.1f}; rounds/group {rounds:.0f}; cycles per round {tot / rounds:.0f}")
m = a.mean(axis=(0, 1))
for k in range(6):
    print(f"   per round: {names[k]:18s} {m[k] / rounds:8.1f}")
