timeout 400 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -6
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 200 --warmup 20 2> gpurun_out/b2.err | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['ms_per_step'], d['roofline']['us_per_launch'], d['config']['launch'], d['parity'])"
tail -3 gpurun_out/b2.err
