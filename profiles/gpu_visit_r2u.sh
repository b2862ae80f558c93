#!/bin/bash
# r2u: lane8 fp32 kernel with the ex2/rcp sigmoid -- fp32 parity tests, full-batch fp32 throughput, fix-up cost
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2u_pytest_gpu.txt
python bench.py --steps 5 --warmup 3 --precision fp32 --queries 4194304 --no-cpu --no-e2e --no-extra > gpurun_out/r2u_bench_fp32_disk.json 2> gpurun_out/r2u_bench.err
python bench.py --steps 5 --warmup 3 --precision fp32 --queries 4194304 --workload spherical --no-cpu --no-e2e --no-extra > gpurun_out/r2u_bench_fp32_spherical.json 2>> gpurun_out/r2u_bench.err
tail -2 gpurun_out/r2u_pytest_gpu.txt
python - <<PY
import json
for w in ("disk","spherical"):
    d=json.load(open(f"gpurun_out/r2u_bench_fp32_{w}.json")); print(w, d["value"], d["ms_per_step"])
PY
