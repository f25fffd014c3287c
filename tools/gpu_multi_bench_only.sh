#!/bin/bash
# the bench line at N GPUs (cells included), nothing else
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 100 --warmup 10 > gpurun_out/r2_bench_n${N}_final.json 2> gpurun_out/r2_bench_n${N}_final.err; echo "bench rc=$?"
tail -2 gpurun_out/r2_bench_n${N}_final.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_n${N}_final.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, d["e2e"]["ms_per_step"], d["roofline"]["us_per_launch"])
for c in d["cells"]:
    print(c["cell"], c.get("us"), c.get("achieved"), c.get("frac"), c.get("error"))
PY
