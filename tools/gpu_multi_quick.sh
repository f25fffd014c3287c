#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-4}
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "8192 or mixed" > gpurun_out/r2_pytest_multi3.log 2>&1; echo "pytest multi rc=$?"
tail -3 gpurun_out/r2_pytest_multi3.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/r2_bench_n${N}c.json 2> gpurun_out/r2_bench_n${N}c.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_bench_n${N}c.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_n${N}c.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, d["e2e"]["ms_per_step"])
for c in d["cells"]:
    print(c["cell"], c.get("us"), c.get("achieved"), c.get("frac"), c.get("error"))
PY
