#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
{
for shape in "q4_k 64128 4096" "q4_k 32064 4096" "q4_k 16032 4096" "q4_k 14336 4096" "q4_k 4096 4096" "q4_k 7168 8192"; do
  for w in 0 1; do
    echo -n "wide=$w "; GGQ_WIDE=$w GGQ_WIDE_MIN_ITEMS=0 timeout 60 python tools/dev_time.py $shape 1 2 2>&1 | tail -1
  done
done
} > gpurun_out/r2_wide2.log 2>&1
cat gpurun_out/r2_wide2.log
