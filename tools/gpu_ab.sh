#!/bin/bash
# dev: A/B timing of build variants (tools/build_variant.sh): gpu_ab.sh "<variants>" "<fmt O K>" ... (T = 1 and 8; also the
# streaming probe GGQ_DECODE_NOCOMPUTE=1 at T = 1)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
variants=$1; shift
{
for shape in "$@"; do
  for v in $variants; do
    if [ $v = base ]; then unset GGQ_LIB_DIR; else export GGQ_LIB_DIR=$PWD/build_variants/$v; fi
    for T in 1 8; do echo -n "$v "; timeout 60 python tools/dev_time.py $shape $T 2 2>&1 | tail -1; done
    echo -n "$v NOCOMPUTE "; GGQ_DECODE_NOCOMPUTE=1 timeout 60 python tools/dev_time.py $shape 1 2 2>&1 | tail -1
  done
done
} > gpurun_out/r2_ab.log 2>&1
cat gpurun_out/r2_ab.log
