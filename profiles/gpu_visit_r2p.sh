#!/bin/bash
# r2p: lane8 fp32 kernel with f32x2 FMAs -- parity suite, fix-up cost per material, full-batch fp32 throughput
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2p_pytest_gpu.txt
python profiles/fixup_cost.py > gpurun_out/r2p_fixup_cost.txt 2> gpurun_out/r2p_fixup_cost.err
python bench.py --steps 5 --warmup 3 --precision fp32 --queries 4194304 --no-cpu --no-e2e --no-extra > gpurun_out/r2p_bench_fp32_disk.json 2> gpurun_out/r2p_bench.err
python bench.py --steps 5 --warmup 3 --precision fp32 --queries 4194304 --workload spherical --no-cpu --no-e2e --no-extra > gpurun_out/r2p_bench_fp32_spherical.json 2>> gpurun_out/r2p_bench.err
tail -2 gpurun_out/r2p_pytest_gpu.txt; cat gpurun_out/r2p_fixup_cost.txt; cat gpurun_out/r2p_bench_fp32_disk.json gpurun_out/r2p_bench_fp32_spherical.json
