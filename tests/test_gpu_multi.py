"""N-split across 2 GPUs (NCCL all-gather path and the fused peer-store path) vs the single-GPU result.
Needs >= 2 GPUs: run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu`."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _worker(rank, world, port, fmt, O, T, K, mode, q):
    try:
        import torch.distributed as td
        for p in (ROOT, os.path.join(ROOT, "gguf-triton-kernel_b200")):
            sys.path.insert(0, p)
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        torch.cuda.set_device(rank)
        td.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        from kernels import _ext
        from multigpu import nsplit
        from oracle import ggq_oracle as orc

        A = torch.from_numpy(orc.random_blocks(fmt, O, K, seed=3)).cuda()
        Xh = np.random.default_rng(4).standard_normal((T, K)).astype(np.float16)
        X = torch.from_numpy(Xh).cuda() if rank == 0 else torch.zeros((T, K), dtype=torch.float16, device="cuda")
        layer = nsplit.NSplitLinear(fmt, nsplit.shard_packed(fmt, A, O, K, world, rank).clone(), O, K, mode=mode,
                                    max_tokens=max(T, 16))
        C = None
        for _ in range(3):  # repeated calls exercise buffer reuse + barriers
            C = layer.forward(X).clone()
        torch.cuda.synchronize()
        full = _ext.mm(_ext.FMT_ID[fmt], A, torch.from_numpy(Xh).cuda(), O, T, K)
        torch.cuda.synchronize()
        diff = (C.float() - full.float()).abs().max().item()
        scale = full.float().abs().max().item()
        ok = C.shape == (T, O) and diff <= 2e-3 * scale
        if mode == "fused" and T <= 16:
            # the fused decode step replayed from a CUDA graph (kernel-maintained epoch), then an eager step again
            if rank == 0:
                layer.set_resident_input(X)
            torch.cuda.synchronize()
            td.barrier()
            replay = layer.capture_steps(T, 4)
            for _ in range(3):
                replay()
            torch.cuda.synchronize()
            ok = ok and torch.equal(layer.last_output(T), C)
            ok = ok and torch.equal(layer.forward(X), C)
            torch.cuda.synchronize()
        q.put((rank, ok, diff, scale))
        td.destroy_process_group()
    except Exception as e:  # noqa: BLE001
        q.put((rank, False, repr(e), 0.0))


@pytest.mark.parametrize("mode", ["nccl", "fused"])
@pytest.mark.parametrize("fmt,O,T,K", [("q4_k", 4096, 1, 4096), ("q6_k", 2048, 8, 2048), ("q8_0", 1024, 256, 1024),
                                        ("q4_k", 8192, 512, 2048)])   # the last: slices >= 1 MB, pushed by the copy engines
def test_nsplit_two_gpus(mode, fmt, O, T, K):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() * 13 + O + T) % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, fmt, O, T, K, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    try:
        res = [q.get(timeout=180) for _ in procs]
        for p in procs:
            p.join(timeout=60)
    finally:
        for p in procs:   # never leave a rank spinning on the GPU
            if p.is_alive():
                p.kill()
    assert all(r[1] is True for r in res), res


def _worker_mixed(rank, world, port, q):
    """a T > 16 GEMM step followed by fused decode steps on the SAME layer (buffer-0 reuse across the two
    exchange disciplines), and 9..16 tokens through the GEMM path."""
    try:
        import torch.distributed as td
        for p in (ROOT, os.path.join(ROOT, "gguf-triton-kernel_b200")):
            sys.path.insert(0, p)
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        torch.cuda.set_device(rank)
        td.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        from kernels import _ext
        from multigpu import nsplit
        from oracle import ggq_oracle as orc
        fmt, O, K = "q4_k", 4096, 2048
        A = torch.from_numpy(orc.random_blocks(fmt, O, K, seed=5)).cuda()
        layer = nsplit.NSplitLinear(fmt, nsplit.shard_packed(fmt, A, O, K, world, rank).clone(), O, K, mode="fused", max_tokens=256)
        ok = True
        worst = 0.0
        for it, T in enumerate([1, 256, 1, 1, 12, 1, 256, 8, 16, 1]):
            Xh = np.random.default_rng(10 + it).standard_normal((T, K)).astype(np.float16)
            X = torch.from_numpy(Xh).cuda()   # same values on every rank; rank 0's are the ones that count
            C = layer.forward(X if rank == 0 else torch.zeros_like(X)).clone()
            full = _ext.mm(_ext.FMT_ID[fmt], A, X, O, T, K)
            torch.cuda.synchronize()
            rel = (C.float() - full.float()).abs().max().item() / full.float().abs().max().item()
            worst = max(worst, rel)
            ok = ok and C.shape == (T, O) and rel <= 2e-3
        ok = ok and layer.sync_status() == 0
        q.put((rank, ok, worst, 0.0))
        td.destroy_process_group()
    except Exception as e:  # noqa: BLE001
        q.put((rank, False, repr(e), 0.0))


def _worker_missing(rank, world, port, q):
    """A rank that never makes the matching call: the waiting rank's kernel gives up after the timeout and reports it."""
    try:
        import time
        import torch.distributed as td
        for p in (ROOT, os.path.join(ROOT, "gguf-triton-kernel_b200")):
            sys.path.insert(0, p)
        os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
        torch.cuda.set_device(rank)
        td.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
        from multigpu import nsplit
        from oracle import ggq_oracle as orc
        fmt, O, K = "q8_0", 2048, 1024
        A = torch.from_numpy(orc.random_blocks(fmt, O, K, seed=6)).cuda()
        layer = nsplit.NSplitLinear(fmt, nsplit.shard_packed(fmt, A, O, K, world, rank).clone(), O, K, mode="fused",
                                    sync_timeout_s=0.25)
        X = torch.randn((1, K), device="cuda", dtype=torch.float16)
        layer.forward(X)                       # a matched step first: status stays 0
        torch.cuda.synchronize()
        ok = layer.sync_status() == 0
        td.barrier()
        code, dt = 0, 0.0
        if rank == 0:                          # rank 1 does NOT make this call
            t0 = time.perf_counter()
            layer.forward(X)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            code = layer.sync_status()
            ok = ok and code == 2 and dt < 5.0    # GGQ_SYNC_TIMEOUT_PEER, in bounded time
        td.barrier()                           # rank 1 stays alive (its memory is mapped by rank 0) until rank 0 is done
        q.put((rank, ok, code, dt))
        td.destroy_process_group()
    except Exception as e:  # noqa: BLE001
        q.put((rank, False, repr(e), 0.0))


def _run2(target, tag):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() * 13 + tag) % 2000
    procs = [ctx.Process(target=target, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    try:
        res = [q.get(timeout=180) for _ in procs]
        for p in procs:
            p.join(timeout=60)
    finally:
        for p in procs:
            if p.is_alive():
                p.kill()
    assert all(r[1] is True for r in res), res


def test_nsplit_fused_mixed_gemm_and_decode_steps():
    _run2(_worker_mixed, 7)


def test_nsplit_missing_rank_times_out_with_status():
    _run2(_worker_missing, 11)
