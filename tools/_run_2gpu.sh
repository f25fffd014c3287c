timeout 300 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -4
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 2> gpurun_out/b2.err | tail -c 1800
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 100 --warmup 10 --exchange nccl 2> gpurun_out/b2n.err | tail -c 600
