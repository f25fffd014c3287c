#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "q4_k" 2>&1 | tail -2
for shape in "q4_k 128256 4096" "q4_k 28672 8192" "q4_k 8192 28672" "q4_k 57344 8192"; do
  for w in 0 8 10; do
    echo -n "wide_nw=$w "; GGQ_WIDE_NW=$w timeout 60 python tools/dev_time.py $shape 1 2 2>&1 | tail -1
  done
done
} > gpurun_out/r2_wide.log 2>&1
cat gpurun_out/r2_wide.log
