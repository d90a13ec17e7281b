#!/bin/bash
# Dev script: build an experimental copy of the product library with extra nvcc flags.
#   tools/build_variant.sh <name> <extra nvcc flags...>   ->  ataraxia_b200/lib/variants/<name>.so   (use with ATX_LIB=...)
set -e
name=$1; shift
cd "$(dirname "$0")/.."
out=ataraxia_b200/lib/variants; mkdir -p $out/obj_$name
for f in atx_capi atx_kernels atx_wavefront atx_p2p; do
  nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -ftz=true -prec-div=false -prec-sqrt=false "$@" \
       -Xcompiler -fPIC,-ffp-contract=off,-fno-fast-math,-fvisibility=hidden -Iinclude -c ataraxia_b200/csrc/$f.cu -o $out/obj_$name/$f.o &
done
wait
nvcc -shared -gencode arch=compute_100a,code=sm_100a $out/obj_$name/*.o -o $out/$name.so -ldl
echo $out/$name.so
