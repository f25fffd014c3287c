#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused_ops.py -m gpu -x -q > gpurun_out/r2_pytest_fused.log 2>&1; echo "pytest fused rc=$?"
tail -5 gpurun_out/r2_pytest_fused.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_all.log 2>&1; echo "pytest all rc=$?"
tail -8 gpurun_out/r2_pytest_all.log
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_bench3.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench3.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","clocks")}, d["e2e"]["ms_per_step"])
for c in d["cells"]:
    print(c["cell"], c["us"], c["achieved"], c["frac"], c.get("frac_8TBps", c.get("frac_2250_nominal")), c["parity"]["ok"])
PY
