for lib in bsdf_diffusion_sampling_b200/libbsdfdiff.so variants/lib_l8u1.so variants/lib_l8u4.so variants/lib_l8u8.so; do
  for w in disk spherical; do
    BSDFDIFF_LIB=$PWD/$lib python bench.py --steps 5 --warmup 3 --precision fp32 --queries 4194304 --workload $w --no-cpu --no-e2e --no-extra 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$lib', '$w', d['value'], d['ms_per_step'])"
  done
done
