#!/usr/bin/env python
"""SASS evidence for the shipped library (no GPU needed: cuobjdump reads the cubin inside libbsdfdiff.so):
    python profiles/sass_histogram.py [lib.so] > profiles/<round>_sass_histogram.txt
Per kernel: instruction count, the tcgen05 / TMA / TMEM opcodes that prove the tensor-core path (UTCHMMA = tcgen05.mma,
LDTM / STTM = tcgen05.ld / st, UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit, SYNCS = mbarrier), MUFU.TANH, and the
top of the opcode histogram.  The activation-pass budget of the hot kernel comes from profiles/sass_pass_count.py."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "bsdf_diffusion_sampling_b200/libbsdfdiff.so"
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", out)), capture_output=True, text=True).stdout.split("\n")
blocks = out.split("Function : ")[1:]
PROOF = ("UTCHMMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "SYNCS", "MUFU.TANH", "ELECT", "R2UR", "FFMA2", "HFMA2", "F2FP")
print(f"# {lib}: {len(blocks)} kernels (cuobjdump -sass, sm_100a)")
total = collections.Counter()
for name, body in zip(names, blocks):
    ops = []
    for line in body.splitlines():
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m:
            ops.append(m.group(3))
    fam = collections.Counter()
    for o in ops:
        key = next((p for p in PROOF if o.startswith(p)), o.split(".")[0])
        fam[key] += 1
        total[key] += 1
    short = re.sub(r"\(bsdfdiff::FlowParams\)|bsdfdiff::|void ", "", name)
    proof = "  ".join(f"{p} {fam[p]}" for p in PROOF if fam[p])
    print(f"\n{short}\n  {len(ops)} instructions | {proof}")
    print("  top: " + ", ".join(f"{k} {v}" for k, v in fam.most_common(12)))
print("\n# whole library: " + "  ".join(f"{p} {total[p]}" for p in PROOF))
