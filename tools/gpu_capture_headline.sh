#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches.csv python bench.py --steps 20 --warmup 5 --no-cells > gpurun_out/r2_ncu_launch.log 2>&1; echo "launch list rc=$?"
cap() {
  local name=$1 rx=$2 skip=$3 script=$4; shift 4
  env "$@" timeout 300 ncu --set full --clock-control none --import-source on -k regex:$rx -c 1 -s $skip -o /tmp/$name -f python $script > gpurun_out/$name.log 2>&1; echo "$name rc=$?"
  python tools/ncu_summary.py /tmp/$name.ncu-rep gpurun_out/${name}_ncu_summary.csv
}
cap r2_decode_q4k_T1_lmhead decode_kernel 3 tools/ncu_one.py FMT=q4_k O=128256 K=4096 T=1
cap r2_swiglu_q4k_T1 decode_kernel 3 tools/ncu_swiglu.py FMT=q4_k O=28672 K=8192 T=1
cap r2_refmode_q4k_T1_lmhead refmode 2 tools/ncu_ref.py FMT=q4_k O=128256 K=4096 T=1
