#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -k "aliasing or multi_material or sharded" 2>&1 | tail -40 > gpurun_out/r2d_pytest_sel.txt
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2d_bench_tc16_disk.json 2> gpurun_out/r2d_bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/r2d_bench_reference.json 2>> gpurun_out/r2d_bench.err
tail -25 gpurun_out/r2d_pytest_sel.txt; cat gpurun_out/r2d_bench_tc16_disk.json gpurun_out/r2d_bench_reference.json; tail -5 gpurun_out/r2d_bench.err
