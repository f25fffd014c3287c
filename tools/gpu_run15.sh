#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
{
for shape in "q4_k 128256 4096" "q6_k 128256 4096" "q8_0 28672 8192" "q4_k 28672 8192" "q4_k 14336 4096" "q8_0 4096 4096"; do
  for T in 1 8; do
    for ah in 0 1; do
      export GGQ_DECODE_L2_AHEAD=$ah
      echo -n "ahead=$ah "; timeout 60 python tools/dev_time.py $shape $T 2 2>&1 | tail -1
    done
  done
done
} > gpurun_out/r2_l2ahead.log 2>&1
cat gpurun_out/r2_l2ahead.log
