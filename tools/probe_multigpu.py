"""torchrun probe: cost of the exchange primitives around the N-split decode step (dev tool)."""
import os, sys, torch, torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gguf-triton-kernel_b200"))
import bench
from kernels import _ext as ext
import torch.distributed._symmetric_memory as symm_mem
rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"]); local = int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
O, K, T = 128256, 4096, 1
rows = O // world
W = bench.gen_weights(torch, "q4_k", rows, K, "cuda", 1 + rank)
x = torch.randn((T, K), device="cuda", dtype=torch.float16)
c = torch.empty((T, rows), device="cuda", dtype=torch.float16)
g = torch.empty((world * T, rows), device="cuda", dtype=torch.float16)
buf = symm_mem.empty((16, O), dtype=torch.float16, device="cuda")
h = symm_mem.rendezvous(buf, dist.group.WORLD)
def t(fn, n=200):
    return bench.timed(torch, dist, fn, n, 20, world) * 1e3
res = {
    "kernel_us": t(lambda: ext.mm(1, W, x, rows, T, K, out=c)),
    "nccl_broadcast_us": t(lambda: dist.broadcast(x, src=0)),
    "nccl_allgather_us": t(lambda: dist.all_gather_into_tensor(g, c)),
    "symm_barrier_us": t(lambda: h.barrier(channel=0)),
}
ptrs = [int(p) + rank * rows * 2 for p in h.buffer_ptrs]
res["kernel_peer_stores_us"] = t(lambda: ext.mm_ex(1, W, x, ptrs, O, rows, T, K))
if rank == 0:
    print(world, res)
dist.destroy_process_group()
