#!/bin/bash
# Tuning build of libbsdfdiff.so with extra -D flags:  profiles/build_variant.sh <name> [-DFLAG ...]  -> variants/lib_<name>.so
# (compared on the GPU box with profiles/variant_compare.py; BSDFDIFF_LIB selects the library).  Same flags as
# bsdf_diffusion_sampling_b200/build.py, including the per-file ones of flow_tc.cu.
set -e
name=$1; shift
cd "$(dirname "$0")/.."
mkdir -p variants/obj_$name
C=bsdf_diffusion_sampling_b200/csrc
for f in capi flow_simt flow_tc measured multi train flow_lane8; do
  extra=""
  if [ $f = flow_tc ]; then extra="-ftz=true -prec-div=false -prec-sqrt=false"; fi
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC \
      --expt-relaxed-constexpr $extra "$@" -c $C/$f.cu -o variants/obj_$name/$f.o 2> >(grep -v "deprecated\|Support for offline" >&2) &
done
wait
/usr/local/cuda/bin/nvcc -shared -o variants/lib_$name.so variants/obj_$name/*.o -lcudart
rm -rf variants/obj_$name
echo variants/lib_$name.so
