#!/usr/bin/env python
"""Static instruction budget of the tensor-core kernel's activation pass, from the SASS of a built library (no GPU needed):
    python profiles/sass_pass_count.py [lib.so] [kernel-substring]
The pass is the straight-line region between the mbarrier wait that precedes the first MUFU.TANH and the group barrier
(BAR.SYNC / BAR.ARV) that follows the last one.  Prints the opcode histogram of that region."""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "bsdf_diffusion_sampling_b200/libbsdfdiff.so"
kern = sys.argv[2] if len(sys.argv) > 2 else "flow_tc_kernelILi0ELi0ELi1ELi32E"     # disk, sample, tanh, H = 32
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
blocks = out.split("Function : ")
body = next(b for b in blocks if kern in b.split("\n", 1)[0])
ins = []
for line in body.splitlines():
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
    if m:
        ins.append(m.group(3))
tanh = [i for i, op in enumerate(ins) if op.startswith("MUFU.TANH")]
tanh = tanh[-32:]                                  # the worker pass (the producer's base net has 16 of its own)
first, last = tanh[0], tanh[-1]
start = max(i for i in range(first) if ins[i].startswith("SYNCS.PHASECHK"))
while start > 0 and not ins[start - 1].startswith(("BRA", "BSYNC")):
    start -= 1
end = min(i for i in range(last, len(ins)) if ins[i].startswith(("BAR.", "ATOMS", "ATOM")))
region = ins[start:end + 1]
hist = collections.Counter(re.sub(r"\..*", "", op) if not op.startswith(("MUFU", "LDTM", "STTM")) else op.split(".PACK")[0] for op in region)
print(f"{lib}  {kern}: {len(region)} instructions in the activation pass ({len(tanh)} MUFU.TANH in the kernel)")
for op, c in sorted(hist.items(), key=lambda kv: -kv[1]):
    print(f"  {op:14s} {c}")
