#!/usr/bin/env python
"""Cycle-stamped timeline of the tensor-core kernel's rounds (tuning build -DBSDFDIFF_TC_TRACE, GPU box):

    BSDFDIFF_LIB=variants/lib_trace.so python profiles/trace_timeline.py [disk|spherical] > gpurun_out/trace.txt

Runs the bench workload once, reads the %clock64 stamps lane 0 of every worker warp of CTA 0 took in 96 steady-state
rounds, and prints (a) the mean / p10 / p90 duration of every segment of a round, per warp position q (= scheduler),
(b) per scheduler, how the 4 worker warps' states overlap in time (share of cycles with k warps inside their
activation pass), (c) the raw stamps of one group for eight rounds."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import bsdf_diffusion_sampling_b200 as pkg  # noqa: E402

W, R, E = 24, 96, 8
SEG = [("wait MMA (prev issue-or-barrier -> wake)", None), ("ld first half (0->1)", (0, 1)),
       ("first half math+st (1->2)", (1, 2)), ("second half math+st (2->3)", (2, 3)),
       ("wait::st + fence (3->4)", (3, 4)), ("group barrier (4->5)", (4, 5)), ("MMA issue, warp 0 (5->6)", (5, 6))]


def main():
    workload = sys.argv[1] if len(sys.argv) > 1 else "disk"
    T = 4 if workload == "disk" else 8
    layers, base = bench.load_fixture(workload)
    pf = pkg.weights.pack_flow_layers(layers, "cuda")
    pb = pkg.weights.pack_base_arrays(*base, "cuda")
    wi = torch.from_numpy(bench.synth_wi3(workload, 4096, 0)).cuda()
    s = pkg.plugins.NeuralBSDFSampler(workload, pf, pb, T=T, precision="tc16", fixup=0.0)
    s.sample(wi, seed=1)
    torch.cuda.synchronize()
    buf = (ctypes.c_ulonglong * (W * R * E))()
    n = pkg._lib.lib.bsdfdiff_debug_trace(ctypes.cast(buf, ctypes.c_void_p), W * R * E)
    if n <= 0:
        print("this library was not built with -DBSDFDIFF_TC_TRACE")
        return
    t = np.frombuffer(buf, dtype=np.uint64).reshape(W, R, E).astype(np.int64)
    t = t[:16]                                                        # worker warps: group g = w // 4, q = w % 4
    t0 = t[:, :, 0].min()
    print(f"workload {workload}: {R} rounds of CTA 0, cycles relative to {t0}")
    period = (t[:, -1, 0] - t[:, 0, 0]) / (R - 1)
    print("round period per warp (cycles):", np.round(period.reshape(4, 4), 1).tolist(), " mean", round(float(period.mean()), 1))
    for name, ev in SEG:
        if ev is None:
            prev_end = np.where(t[:, :-1, 6] > 0, t[:, :-1, 6], t[:, :-1, 5])
            d = t[:, 1:, 0] - prev_end
        else:
            d = t[:, :, ev[1]] - t[:, :, ev[0]]
            if ev == (5, 6):
                d = d[0::4]
        rows = []
        for q in range(4 if ev != (5, 6) else 1):
            x = d[q::4].ravel() if ev != (5, 6) else d.ravel()
            rows.append(f"q{q}: mean {x.mean():6.0f} p10 {np.quantile(x, .1):6.0f} p90 {np.quantile(x, .9):6.0f}")
        print(f"{name:44s} " + " | ".join(rows))
    # overlap per scheduler: warps w with w % 4 == q share scheduler q
    lo, hi = int(t[:, 8, 0].max()), int(t[:, -8, 5].min())
    for q in range(4):
        cnt = np.zeros(hi - lo, np.int32)
        ld = np.zeros(hi - lo, np.int32)
        for w in range(q, 16, 4):
            for r in range(R):
                a, b = t[w, r, 0] - lo, t[w, r, 4] - lo
                cnt[max(a, 0):max(min(b, hi - lo), 0)] += 1
        share = [float((cnt == k).mean()) for k in range(5)]
        print(f"scheduler {q}: share of cycles with k worker warps inside an activation pass (wake -> stores landed), k=0..4: "
              + " ".join(f"{x:.3f}" for x in share) + f"   mean k = {cnt.mean():.2f}")
    g = 1
    print(f"raw stamps, group {g}, rounds 20..27 (relative cycles; events 0..6):")
    for r in range(20, 28):
        for q in range(4):
            print(f"  r{r} q{q}: " + " ".join(f"{int(v - t0):8d}" for v in t[4 * g + q, r, :7]))


if __name__ == "__main__":
    main()
