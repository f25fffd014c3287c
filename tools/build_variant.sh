#!/bin/bash
# dev helper: build_variants/libggq_<name>.so with extra -D flags (A/B experiments on the GPU box)
set -e
name=$1; shift
cd "$(dirname "$0")/../gguf-triton-kernel_b200"
out=../build_variants/$name
mkdir -p $out
for f in api generic decode prefill pack refmode; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c csrc/$f.cu -o $out/$f.o &
done
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o ../build_variants/libggq_$name.so $out/*.o
rm -rf $out
