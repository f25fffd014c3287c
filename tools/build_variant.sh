#!/bin/bash
# dev helper: build_variants/<name>/{libggq.so,_ggq_torch.so} with extra -D flags (A/B experiments on the GPU box:
# GGQ_LIB_DIR=build_variants/<name> selects it in kernels/_ext.py)
set -e
name=$1; shift
cd "$(dirname "$0")/../gguf-triton-kernel_b200"
out=../build_variants/$name
mkdir -p $out/obj
for f in api generic decode decode_dual prefill skinny pack refmode host; do
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr "$@" -c csrc/$f.cu -o $out/obj/$f.o &
done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr -fmad=false "$@" -c csrc/packk.cu -o $out/obj/packk.o &
wait
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared -o $out/libggq.so $out/obj/*.o
cp _ggq_torch.so $out/   # links libggq.so through an $ORIGIN rpath: picks the variant next to it
rm -rf $out/obj
