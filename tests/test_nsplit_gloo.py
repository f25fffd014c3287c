"""World-size-2 gloo test (CPU) of the N-split host logic: sharding arithmetic, broadcast, all-gather
layout.  The per-rank matmul is a stand-in (the oracle's fp32 reference) because the CUDA kernels need a
GPU; the GPU path with NCCL / peer stores is covered by tests/test_gpu_multi.py and bench.py --gpus N."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, fmt, O, T, K, q):
    import torch.distributed as td
    for p in (ROOT, os.path.join(ROOT, "gguf-triton-kernel_b200")):
        sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    td.init_process_group("gloo", rank=rank, world_size=world)
    from multigpu import nsplit
    from oracle import ggq_oracle as orc

    A = torch.from_numpy(orc.random_blocks(fmt, O, K, seed=3))           # same on every rank
    X = torch.from_numpy(np.random.default_rng(4).standard_normal((T, K)).astype(np.float16))
    Xr = X.clone() if rank == 0 else torch.zeros_like(X)                   # only rank 0 has the activations

    def mm_fn(a, x, m, n, k):  # stand-in for the CUDA kernel
        return torch.from_numpy(orc.ref32(fmt, a.numpy(), x.numpy(), m, n, k).astype(np.float16))

    layer = nsplit.NSplitLinear(fmt, nsplit.shard_packed(fmt, A, O, K, world, rank), O, K, mm_fn=mm_fn)
    C = layer.forward(Xr)
    want = orc.ref32(fmt, A.numpy(), X.numpy(), O, T, K).astype(np.float16)
    ok = C.shape == (T, O) and np.array_equal(C.numpy().view(np.uint16), want.view(np.uint16))
    q.put((rank, bool(ok)))
    td.destroy_process_group()


@pytest.mark.parametrize("fmt,O,T,K", [("q4_k", 64, 1, 512), ("q6_k", 32, 5, 256), ("q8_0", 48, 16, 128)])
def test_nsplit_world2_gloo(fmt, O, T, K):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() * 7 + O) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, fmt, O, T, K, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)], res


def test_shard_arithmetic():
    sys.path.insert(0, os.path.join(ROOT, "gguf-triton-kernel_b200"))
    from multigpu import nsplit
    assert nsplit.row_bytes("q4_k", 4096) == 2304 and nsplit.row_bytes("q6_k", 4096) == 3360
    assert nsplit.shard_rows(128256, 8, 7) == (112224, 128256)
    with pytest.raises(ValueError):
        nsplit.shard_rows(10, 4, 0)
    g = torch.arange(2 * 3 * 4).reshape(2, 3, 4)
    full = nsplit.assemble(g)
    assert full.shape == (3, 8) and full[1].tolist() == [4, 5, 6, 7, 16, 17, 18, 19]
