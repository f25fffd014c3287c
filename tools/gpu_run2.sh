#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 300 python tools/probe_e2e.py > gpurun_out/r2_probe_e2e.log 2>&1; echo "probe rc=$?"
cat gpurun_out/r2_probe_e2e.log
timeout 200 python tools/dev_skinny.py parity 2>&1 | grep -v "^ok\|^skip" | tail -5
for T in 16 8; do
timeout 60 python tools/dev_time.py q4_k 128256 4096 $T 4
timeout 60 python tools/dev_time.py q6_k 128256 4096 $T 4
timeout 60 python tools/dev_time.py q8_0 28672 8192 $T 4
done
timeout 60 python tools/dev_time.py q4_k 14336 4096 16 4
