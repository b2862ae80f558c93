#!/usr/bin/env python
"""Throughput of the shipped path on EVERY material the reference ships (oracle/_ref/all_*.bsdfpack): sample() of 4.19 M queries
with the flow family's default fix-up thresholds and with the material's own calibrated thresholds, against the tensor-core
launch alone.  Shows what the fix-up pass costs across the real material set (the bench material is one of the cheap ones).
    python profiles/material_perf.py > profiles/<round>_material_perf.txt"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bsdf_diffusion_sampling_b200 as pkg          # noqa: E402
from bsdf_diffusion_sampling_b200.materials import MaterialPack   # noqa: E402


def timed(fn, steps=3, warm=1):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / steps


n_side = 2048
print("%-10s %-36s %9s | %10s %8s | %10s %8s %-14s" % ("kind", "material", "tc only", "default", "fixed %", "calibrated", "fixed %", "thresholds s/p"))
summary = {}
for kind in ("disk", "spherical", "bsdf"):
    pack = MaterialPack.load(os.path.join(ROOT, "oracle", "_ref", f"all_{kind}.bsdfpack"))
    wi = torch.from_numpy(pkg.plugins.stratified_domain_wi(kind, n_side, 5)).cuda()
    n = wi.shape[0]
    for e in pack.entries:
        pf = pkg.weights.pack_flow_layers(e["flow"], "cuda")
        pb = torch.from_numpy(np.array(e["base"])).cuda()
        T = e["T"]
        fam = pkg.ops._fix_thr(None, pf.domain, pkg.plugins._KINDS[kind][1], "sample")
        cal = pkg.plugins.NeuralBSDFSampler(kind, pf, pb, T=T).calibrate_fixup(n_side=128, seed=7, install=False)
        t0 = timed(lambda: pkg.ops.sample(wi, pf, pb, T, seed=3, precision="tc16", fixup=0.0, return_x0=False))
        t1 = timed(lambda: pkg.ops.sample(wi, pf, pb, T, seed=3, precision="tc16", fixup=fam, return_x0=False))
        f1 = pkg.ops.last_fixup_count() / n
        t2 = timed(lambda: pkg.ops.sample(wi, pf, pb, T, seed=3, precision="tc16", fixup=cal["sample"], return_x0=False))
        f2 = pkg.ops.last_fixup_count() / n if cal["sample"] > 0 else 0.0
        summary.setdefault(kind, []).append((n / t0 / 1e6, n / t1 / 1e6, n / t2 / 1e6))
        print("%-10s %-36s %8.2fe9 | %9.2fe9 %8.3f | %9.2fe9 %8.3f %.4g/%.4g" % (
            kind, e["name"], n / t0 / 1e6, n / t1 / 1e6, 100 * f1, n / t2 / 1e6, 100 * f2, cal["sample"], cal["pdf"]), flush=True)
for kind, v in summary.items():
    a = np.array(v)
    print(f"{kind}: sample() q/s over {len(v)} materials -- tensor-core launch alone median {np.median(a[:, 0]):.2f}e9; family-default "
          f"thresholds median {np.median(a[:, 1]):.2f}e9, mean {a[:, 1].mean():.2f}e9, min {a[:, 1].min():.2f}e9; per-material calibrated "
          f"median {np.median(a[:, 2]):.2f}e9, mean {a[:, 2].mean():.2f}e9, min {a[:, 2].min():.2f}e9")
