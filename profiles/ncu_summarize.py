#!/usr/bin/env python
"""Summarise an .ncu-rep (raw + source pages) into a small text file for profiles/.
usage: python profiles/ncu_summarize.py gpurun_out/x.ncu-rep profiles/x_summary.txt "title" """
import collections
import csv
import io
import re
import subprocess
import sys

rep, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active",
        "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "SM_A.TriageCompute.sm__inst_executed_pipe_xu_realtime.avg.pct_of_peak_sustained_elapsed"]
keep += [h for h in hdr if "issue_stalled" in h and h.endswith("per_issue_active.ratio")]
lines = [title]
for h, u, v in zip(hdr, units, vals):
    if h in keep:
        lines.append(f"{h} = {v} {u}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
h2, data = rows[1], rows[2:]
ia, ie, isamp = h2.index("Source"), h2.index("Instructions Executed"), h2.index("Warp Stall Sampling (All Samples)")
tot_e = sum(int(r[ie]) for r in data)
tot_s = sum(int(r[isamp]) for r in data) or 1
op, ops = collections.Counter(), collections.Counter()
for r in data:
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ia])
    full = m.group(2) if m else "?"
    o = full.split(".")[0]
    key = o if o not in ("MUFU", "F2FP", "SYNCS", "LDTM", "STTM", "BAR") else full
    op[key] += int(r[ie])
    ops[key] += int(r[isamp])
lines.append(f"--- SASS opcode mix (total warp-instructions {tot_e}, stall samples {tot_s})")
lines.append("%-32s %8s %8s" % ("opcode", "exec %", "samples %"))
for k, v in op.most_common(28):
    lines.append("%-32s %8.2f %8.2f" % (k, 100 * v / tot_e, 100 * ops[k] / tot_s))
lines.append("--- top stall sites (samples, executed, SASS)")
for r in sorted(data, key=lambda r: -int(r[isamp]))[:14]:
    lines.append(f"{r[isamp]:>8} {r[ie]:>10}  {r[ia].strip()[:100]}")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines))
