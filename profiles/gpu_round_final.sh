#!/bin/bash
# Final GPU visit of a round: parity suite, bench lines (ours + reference arm + spherical), launch list, ncu --set full of the
# hot kernel, material sweep, training bench, compute-sanitizer memcheck over the GPU suite.
# usage: gpurun --timeout 1500 -- 'bash profiles/gpu_round_final.sh r2w'
tag=${1:-rX}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/${tag}_pytest_gpu.txt
python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench_tc16_disk.json 2> gpurun_out/${tag}_bench_disk.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench_disk.err
python bench.py --steps 10 --warmup 3 --workload spherical --no-cpu --no-extra > gpurun_out/${tag}_bench_tc16_spherical.json 2>> gpurun_out/${tag}_bench_disk.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/${tag}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:flow_tc_kernel -s 3 -c 1 -f -o gpurun_out/${tag}_tc16_disk \
    python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-extra > gpurun_out/${tag}_ncu_full.log 2>&1
python profiles/material_sweep.py 256 > gpurun_out/${tag}_material_sweep.txt 2> gpurun_out/${tag}_material_sweep.err
python profiles/train_bench.py --workload disk 2>/dev/null | tail -1 > gpurun_out/${tag}_train_bench.jsonl
python profiles/train_bench.py --workload spherical 2>/dev/null | tail -1 >> gpurun_out/${tag}_train_bench.jsonl
python profiles/fixup_cost.py > gpurun_out/${tag}_fixup_cost.txt 2>/dev/null
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests -m gpu -x -q -k "not every_shipped and not noalias" 2>&1 | tail -16 > gpurun_out/${tag}_sanitizer_memcheck.txt
tail -2 gpurun_out/${tag}_pytest_gpu.txt; cat gpurun_out/${tag}_bench_tc16_disk.json | cut -c1-600; cat gpurun_out/${tag}_bench_reference.json | cut -c1-400; tail -4 gpurun_out/${tag}_sanitizer_memcheck.txt
