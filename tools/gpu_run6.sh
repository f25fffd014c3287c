#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
{
timeout 120 python tools/trace_decode.py q4_k 128256 4096 1 6
timeout 120 python tools/trace_decode.py q4_k 16032 4096 1 6
timeout 120 python tools/trace_decode.py q8_0 4096 4096 1 6
timeout 120 python tools/trace_decode.py q4_k 14336 4096 1 6
} > gpurun_out/r2_trace.log 2>&1
cat gpurun_out/r2_trace.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "baseline" > gpurun_out/r2_pytest_baseline.log 2>&1; echo "pytest baseline rc=$?"
tail -5 gpurun_out/r2_pytest_baseline.log
