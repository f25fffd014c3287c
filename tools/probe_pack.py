"""dev helper: K-quant packer throughput on the GPU vs the reference library on one host core."""
import sys, os, time, ctypes
import numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/gguf-triton-kernel_b200")
from utils.quantize.q4_k import quantize_to_q4_k
from utils.quantize.q6_k import quantize_to_q6_k
O, K = 14336, 4096
W = torch.randn((O, K), device="cuda", dtype=torch.float32)
for name, fn, lib, sym, blk in (("q4_k", quantize_to_q4_k, "libq4_k_ref.so", "quantize_row_q4_K_ref", 144),
                                ("q6_k", quantize_to_q6_k, "libq6_k_ref.so", "quantize_row_q6_K_ref", 210)):
    fn(W); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3): A = fn(W)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    rows = 256
    x = W[:rows].cpu().numpy().ravel()
    out = np.zeros(x.size // 256 * blk, dtype=np.uint8)
    L = ctypes.CDLL(f"/root/repo/oracle/_ref/{lib}")
    t0 = time.perf_counter()
    getattr(L, sym)(x.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p), ctypes.c_longlong(x.size))
    dt = time.perf_counter() - t0
    same = np.array_equal(A[: out.size].cpu().numpy().view(np.uint8), out)
    print(f"{name}: GPU {ms:.2f} ms for {O}x{K} ({O*K/ms/1e6:.1f} Gweights/s); reference C, 1 core: {dt*O/rows:.1f} s extrapolated from {rows} rows "
          f"({x.size/dt/1e6:.1f} Mweights/s); speed-up {dt*O/rows/(ms/1e3):.0f}x; first {rows} rows identical: {same}")
