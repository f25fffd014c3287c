#!/bin/bash
# full single-GPU validation: every GPU test, smoke(), the bench line
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_final.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2_pytest_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2_smoke.log
timeout 900 python bench.py > gpurun_out/r2_bench_final.json 2> gpurun_out/r2_bench_final.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_bench_final.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2_bench_final.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","clocks","gpu_launches")}, d["e2e"]["ms_per_step"], d["roofline"]["traffic"], d["roofline"]["frac"], d["roofline"]["isolated"])
print(d["roofline"]["kernel"])
for c in d["cells"]:
    print(c["cell"], c["us"], c["achieved"], c["frac"], c.get("frac_8TBps", c.get("frac_2250_nominal")), c["parity"]["ok"])
PY
