#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused_ops.py -m gpu -x -q > gpurun_out/r2_pytest_fused.log 2>&1; echo "pytest fused rc=$?"
tail -15 gpurun_out/r2_pytest_fused.log
timeout 900 python bench.py --steps 100 --warmup 10 > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_bench2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench2.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","e2e","clocks")})
print(d["roofline"])
for c in d["cells"]:
    print(c["cell"], c["us"], c["achieved"], c["frac"], c.get("us_two_mmq_plus_torch_silu_mul"), c["parity"]["ok"])
PY
