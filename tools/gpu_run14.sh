#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
{
for shape in "q8_0 4096 4096" "q4_k 4096 4096" "q4_k 14336 4096" "q6_k 4096 14336" "q4_k 28672 8192" "q4_k 8192 28672"; do
  for T in 1 8; do
    for plan in default 8,2 12,1 8,1; do
      if [ $plan = default ]; then unset GGQ_PLAN_FORCE; else export GGQ_PLAN_FORCE=$plan; fi
      echo -n "plan=$plan "; timeout 60 python tools/dev_time.py $shape $T 2 2>&1 | tail -1
    done
  done
done
} > gpurun_out/r2_plans.log 2>&1
cat gpurun_out/r2_plans.log
