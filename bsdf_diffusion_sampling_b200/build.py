"""Build libbsdfdiff.so (the C-ABI library with every CUDA kernel) in-tree with nvcc for sm_100a.

    python -m bsdf_diffusion_sampling_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU; the resulting .so sits next to this file (git-ignored) and is what
``_lib.py`` loads.  There is no JIT and no fallback: if the library is missing the package raises.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libbsdfdiff.so")
SOURCES = ["capi.cu", "flow_simt.cu", "flow_tc.cu", "measured.cu", "multi.cu", "train.cu", "flow_lane8.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]
# per-source extras.  The tensor-core path works at fp16-operand accuracy, so its fp32 prologue/epilogue math (base
# net, Box-Muller, domain maps) uses flush-to-zero and the approximate divide / square root: ~15 % fewer SASS
# instructions in the producer warps, which share issue slots with the activation math.  The fp32 parity kernel
# (flow_simt.cu) keeps IEEE division and square roots.
EXTRA_FLAGS = {"flow_tc.cu": ["-ftz=true", "-prec-div=false", "-prec-sqrt=false"]}


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, "..", "include", "bsdfdiff.h")]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(CSRC, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *EXTRA_FLAGS.get(src, []), "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError("nvcc failed: " + " ".join(cmd))
    subprocess.check_call([nvcc, "-shared", "-o", LIB, *objs, "-lcudart"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
