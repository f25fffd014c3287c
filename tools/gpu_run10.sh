#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
FMT=q6_k O=128256 K=4096 T=16 timeout 300 ncu --set full --clock-control none --import-source on -k regex:skinny -c 1 -s 3 -o gpurun_out/r2_skinny_q6k_t16 -f python tools/ncu_one.py > gpurun_out/r2_ncu3.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r2_ncu3.log
FMT=q6_k O=128256 K=4096 T=2048 timeout 300 ncu --set full --clock-control none --import-source on -k regex:prefill2 -c 1 -s 2 -o gpurun_out/r2_prefill_q6k -f python tools/ncu_one.py > gpurun_out/r2_ncu4.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/r2_ncu4.log
