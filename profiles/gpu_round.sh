#!/bin/bash
# One GPU-box visit: parity tests, bench lines, launch list and one ncu --set full capture of the hot kernel.
# usage (from the repo root): gpurun --timeout 900 -- 'bash profiles/gpu_round.sh r1m'
tag=${1:-rX}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/${tag}_pytest_gpu.txt
python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench_tc16_disk.json 2> gpurun_out/${tag}_bench_disk.err
python bench.py --steps 10 --warmup 3 --workload spherical --no-cpu --no-e2e > gpurun_out/${tag}_bench_tc16_spherical.json 2>> gpurun_out/${tag}_bench_disk.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/${tag}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:flow_tc_kernel -s 3 -c 1 -f -o gpurun_out/${tag}_tc16_disk \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/${tag}_ncu_full.log 2>&1
tail -2 gpurun_out/${tag}_pytest_gpu.txt; cat gpurun_out/${tag}_bench_tc16_disk.json; cat gpurun_out/${tag}_bench_tc16_spherical.json
