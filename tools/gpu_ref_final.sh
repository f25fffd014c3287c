#!/bin/bash
# every GPU test, then the reference-arithmetic mode's timings and ncu --set full captures (T = 1 tile kernel, T = 8 tensor-core kernel)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_final.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2_pytest_final.log
timeout 120 python tools/dev_ref_time.py 1 2 3 4 8 16 > gpurun_out/r2_ref5_time.log 2>&1; cat gpurun_out/r2_ref5_time.log
cap() {  # name, kernel regex, skip, script, env...
  local name=$1 rx=$2 skip=$3 script=$4; shift 4
  env "$@" timeout 200 ncu --set full --clock-control none --import-source on -k regex:$rx -c 1 -s $skip -o /tmp/$name -f python $script > gpurun_out/$name.log 2>&1; echo "$name rc=$?"
  python tools/ncu_summary.py /tmp/$name.ncu-rep gpurun_out/${name}_ncu_summary.csv
}
cap r2_refmode_tile_q4k_T1_lmhead refmode_tile 2 tools/ncu_ref.py FMT=q4_k O=128256 K=4096 T=1
cap r2_refmode_mma_q4k_T8_lmhead refmode_mma 2 tools/ncu_ref.py FMT=q4_k O=128256 K=4096 T=8
