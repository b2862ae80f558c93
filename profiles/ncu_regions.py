#!/usr/bin/env python
"""Attribute the warp-stall samples of an `ncu --set full --import-source on` capture of flow_tc_kernel to the
kernel's regions, using the per-instruction execution counts as the region key (an instruction executed 12 x
warp-tiles times belongs to the activation round, 4 x to the per-step code, 1 x to the per-tile code, anything above
to the mbarrier wait loops).
usage: python profiles/ncu_regions.py gpurun_out/x.ncu-rep [T=4] [NH=3] > profiles/x_stall_regions.txt"""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
T = int(sys.argv[2]) if len(sys.argv) > 2 else 4
NH = int(sys.argv[3]) if len(sys.argv) > 3 else 3
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h, data = rows[1], rows[2:]
ia, ie, isamp = h.index("Source"), h.index("Instructions Executed"), h.index("Warp Stall Sampling (All Samples)")
reasons = [c for c in h if c.startswith("stall_") and "Not Issued" not in c]
idx = {r: h.index(r) for r in reasons}
tanh = collections.Counter(int(r[ie]) for r in data if "MUFU.TANH" in r[ia])
per_round = tanh.most_common(1)[0][0]                      # executions of an activation-round instruction
per_tile = per_round // (T * NH)
per_step = per_tile * T


def region(e):
    if e == per_round:
        return "activation round (x%d per tile)" % (T * NH)
    if e == per_step:
        return "per Euler step (x%d per tile)" % T
    if abs(e - per_tile) <= per_tile // 100:
        return "per tile (worker prologue/epilogue + producer)"
    if e > per_round * 0.6:
        return "mbarrier wait loops"
    if e == 0:
        return "never executed"
    return "MMA issue (one lane) / other"


agg = collections.defaultdict(collections.Counter)
tot, ninst, nexec = collections.Counter(), collections.Counter(), collections.Counter()
for r in data:
    e = int(r[ie])
    reg = region(e)
    tot[reg] += int(r[isamp])
    ninst[reg] += 1
    nexec[reg] += e
    for k, i in idx.items():
        agg[reg][k[6:]] += int(r[i] or 0)
all_s, all_e = sum(tot.values()), sum(nexec.values())
print(f"{rep}: {all_s} stall samples, {all_e} warp-instructions; warp-tiles = {per_tile}")
for reg, s in tot.most_common():
    if reg == "never executed":
        continue
    print(f"\n{reg}: {ninst[reg]} SASS instructions, {100.0 * nexec[reg] / all_e:.1f} % of executed, "
          f"{100.0 * s / all_s:.1f} % of samples")
    print("   " + "  ".join(f"{k} {100.0 * v / max(s, 1):.0f}%" for k, v in agg[reg].most_common(7)))
