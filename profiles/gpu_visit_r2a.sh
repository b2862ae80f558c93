#!/bin/bash
# round-2 visit A: parity suite with the raw bar + fix-up, error statistics per fix-up threshold, bench line
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q -s 2>&1 | tail -60 > gpurun_out/r2a_pytest_gpu.txt
python profiles/tc_error_stats.py tc16 > gpurun_out/r2a_tc16_error_stats.txt 2>&1
python bench.py --steps 20 --warmup 3 > gpurun_out/r2a_bench_tc16_disk.json 2> gpurun_out/r2a_bench.err
BSDFDIFF_FIXUP=0 python bench.py --steps 20 --warmup 3 --no-cpu --no-e2e > gpurun_out/r2a_bench_tc16_disk_nofix.json 2>> gpurun_out/r2a_bench.err
tail -5 gpurun_out/r2a_pytest_gpu.txt; cat gpurun_out/r2a_bench_tc16_disk.json gpurun_out/r2a_bench_tc16_disk_nofix.json
