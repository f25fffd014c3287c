#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
for T in 256 40; do
FMT=q6_k O=1024 K=2048 T=$T timeout 120 python tools/ncu_one.py > gpurun_out/r2_q6k_T$T.log 2>&1; echo "T=$T rc=$?"; tail -3 gpurun_out/r2_q6k_T$T.log
done
FMT=q6_k O=1024 K=2048 T=40 timeout 300 compute-sanitizer --tool memcheck python tools/ncu_one.py 2>&1 | grep -v "^=========     at\|^=========         in\|Host Frame\|^=========$" | head -40
