#!/bin/bash
# round-2 visit B: A/B of the two-tiles-per-thread ("duo") worker structure against the round-1 structure
mkdir -p gpurun_out
BSDFDIFF_FIXUP=0 timeout 900 python profiles/variant_compare.py bsdf_diffusion_sampling_b200/libbsdfdiff.so variants/lib_duo_i1.so variants/lib_duo_i0.so variants/lib_duo_sk0.so variants/lib_old.so > gpurun_out/r2b_variants.txt 2>&1
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2b_pytest_gpu.txt
cat gpurun_out/r2b_variants.txt | grep -v "_err\|sha1\|pdf_sum\|nonfinite"; tail -4 gpurun_out/r2b_pytest_gpu.txt
