#!/bin/bash
# quick N-GPU regression: the fused-exchange tests + the headline line (no cells)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
N=${1:-2}
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "q4_k" > gpurun_out/r2_pytest_q4k.log 2>&1; echo "pytest q4_k rc=$?"; tail -2 gpurun_out/r2_pytest_q4k.log
timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -k "fused and (4096 or mixed)" > gpurun_out/r2_pytest_multi3.log 2>&1; echo "pytest multi rc=$?"
tail -3 gpurun_out/r2_pytest_multi3.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 100 --warmup 10 --no-cells > gpurun_out/r2_bench_n${N}d.json 2> gpurun_out/r2_bench_n${N}d.err; echo "bench rc=$?"
tail -2 gpurun_out/r2_bench_n${N}d.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_n${N}d.json").read().strip().splitlines()[-1])
print({k:d[k] for k in ("value","ms_per_step","n_gpus")}, d["e2e"]["ms_per_step"], d["roofline"]["us_per_launch"], d["parity"])
PY
