#!/usr/bin/env python
"""What the conditioning-triggered fp32 fix-up costs per material: 16.7 M sample() queries with the bench's synthetic wi,
tensor-core launch alone (fixup=0) vs the shipped path (the flow family's default threshold) and vs the material's own
calibrated threshold (NeuralBSDFSampler.calibrate_fixup), rows recomputed, for every golden material.
    python profiles/fixup_cost.py > profiles/<round>_fixup_cost.txt"""
import glob
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                    # noqa: E402
import bsdf_diffusion_sampling_b200 as pkg      # noqa: E402


def timed(fn, steps=5, warm=2):
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


print("%-44s %5s %10s %10s %9s %-22s | %s" % ("material", "T", "tc only ms", "shipped ms", "overhead", "rows fixed", "calibrated: thr, ms, rows fixed"))
for path in sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz"))):
    z = np.load(path)
    name = os.path.basename(path)[:-4]
    kind = "disk" if name.startswith("disk_") else ("bsdf" if name.startswith("bsdf_") else "spherical")
    pf = pkg.weights.pack_flow_layers([z[f"flow_w{i}"] for i in range(int(z["n_flow_layers"]))], "cuda")
    pb = pkg.weights.pack_base_arrays(z["base_w1"], z["base_b1"], z["base_wo"], z["base_bo"], "cuda")
    wi = torch.from_numpy(bench.synth_wi3("disk" if kind == "disk" else "spherical", 4096, 5)).cuda()
    if kind == "bsdf":
        wi[1::2, 2] *= -1.0                                   # both hemispheres
    a = pkg.plugins.NeuralBSDFSampler(kind, pf, pb, precision="tc16", fixup=0.0)
    b = pkg.plugins.NeuralBSDFSampler(kind, pf, pb, precision="tc16")
    c = pkg.plugins.NeuralBSDFSampler(kind, pf, pb, precision="tc16")
    cal = c.calibrate_fixup()
    ta = timed(lambda: a.sample(wi, seed=3))
    tb = timed(lambda: b.sample(wi, seed=3))
    rows = pkg.ops.last_fixup_count()
    tc = timed(lambda: c.sample(wi, seed=3))
    rows_c = pkg.ops.last_fixup_count() if cal["sample"] > 0 else 0
    print("%-44s %5d %10.3f %10.3f %8.1f%% %9d (%.3f%%)   | %.4g  %.3f ms  %d (%.3f%%)" % (
        name, a.T, ta, tb, 100 * (tb / ta - 1), rows, 100 * rows / wi.shape[0], cal["sample"], tc, rows_c, 100 * rows_c / wi.shape[0]))
