#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "mm_sync or chain or mm_ex" > gpurun_out/r2_pytest_sync.log 2>&1; echo "pytest sync rc=$?"
tail -15 gpurun_out/r2_pytest_sync.log
