#!/bin/bash
mkdir -p gpurun_out
BSDFDIFF_LIB=$PWD/variants/lib_trace.so timeout 300 python profiles/trace_timeline.py disk > gpurun_out/r2c_trace_disk.txt 2>&1
BSDFDIFF_LIB=$PWD/variants/lib_trace.so timeout 300 python profiles/trace_timeline.py spherical > gpurun_out/r2c_trace_spherical.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2c_pytest_gpu.txt
cat gpurun_out/r2c_trace_disk.txt; tail -4 gpurun_out/r2c_pytest_gpu.txt
