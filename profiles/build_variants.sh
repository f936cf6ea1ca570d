#!/bin/bash
# Builds A/B variants of libbamse_cuda.so into variants/<name>/ (git-ignored, travels with gpurun).
# usage: profiles/build_variants.sh name "-DMACRO ..." [name "-D..."]...
# select at run time with BAMSE_CUDA_LIB=variants/<name>/libbamse_cuda.so
set -e
cd "$(dirname "$0")/../bamse_b200/csrc"
FLAGS="-O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC,-O3 --fmad=true"
while [ $# -ge 2 ]; do
    name=$1; defs=$2; shift 2
    out=../../variants/$name; mkdir -p $out
    for f in capi kernels fast; do nvcc $FLAGS $defs -c $f.cu -o $out/$f.o & done; wait
    nvcc -shared -gencode arch=compute_100a,code=sm_100a -o $out/libbamse_cuda.so $out/capi.o $out/kernels.o $out/fast.o
    rm -f $out/*.o
    echo "built variants/$name ($defs)"
done
