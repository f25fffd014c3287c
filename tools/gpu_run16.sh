#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_quantize_ops.py -m gpu -x -q -k "reference_arithmetic" > gpurun_out/r2_pytest_ref.log 2>&1; echo "pytest rc=$?"
tail -6 gpurun_out/r2_pytest_ref.log
timeout 900 python bench.py --steps 50 --warmup 5 > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_bench4.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench4.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","clocks")}, d["e2e"]["ms_per_step"], d["roofline"]["traffic"])
for c in d["cells"]:
    print(c["cell"], c["us"], c["achieved"], c["frac"], c["parity"]["ok"])
PY
