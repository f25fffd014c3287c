#!/bin/bash
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q > gpurun_out/r2_pytest_multi2.log 2>&1; echo "pytest multi rc=$?"
tail -5 gpurun_out/r2_pytest_multi2.log
N=${1:-2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 50 --warmup 5 > gpurun_out/r2_bench_n${N}b.json 2> gpurun_out/r2_bench_n${N}b.err; echo "bench rc=$?"
tail -3 gpurun_out/r2_bench_n${N}b.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_n${N}b.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, d["e2e"]["ms_per_step"])
for c in d["cells"]:
    print(c["cell"], c["us"], c["achieved"], c["frac"], c["parity"].get("ok", c["parity"]))
PY
