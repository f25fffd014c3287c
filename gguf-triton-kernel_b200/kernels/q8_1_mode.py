"""Reference-arithmetic mode on the GPU: the same signatures as the reference's CPU implementations
(kernels/cpu_impls/mmq_q8_0_q8_1_cpu.py:5, mmq_q4_k_q8_1_cpu.py:61, mmq_q6_k_q8_1_cpu.py:84) — Q8_1-packed
activations in, fp16 [N, M] out — and bit-identical results (integer block dots, fp16 accumulator, same
operation order).  For parity work; the fast path is kernels/mmq_*.py."""
import ctypes

import torch

from . import _ext

_ready = False


def _call(fmt: int, A: torch.Tensor, B: torch.Tensor, M: int, N: int, K: int) -> torch.Tensor:
    global _ready
    L = _ext.lib()
    if not _ready:
        L.ggq_mm_ref_q8_1.argtypes = [ctypes.c_int] + [ctypes.c_void_p] * 3 + [ctypes.c_int64] * 3 + [ctypes.c_void_p]
        L.ggq_mm_ref_q8_1.restype = ctypes.c_int
        _ready = True
    assert K % _ext.FMT_QK[fmt] == 0
    assert A.dtype == torch.int8 and B.dtype == torch.int8
    assert A.numel() == M * (K // _ext.FMT_QK[fmt]) * _ext.FMT_BLK[fmt]
    assert B.numel() == N * (K // 32) * 36
    if not (A.is_cuda and B.is_cuda):
        raise ValueError("CUDA tensors expected (no CPU path)")
    C = torch.empty((N, M), dtype=torch.float16, device=A.device)
    with torch.cuda.device(A.device):
        rc = L.ggq_mm_ref_q8_1(fmt, A.contiguous().data_ptr(), B.contiguous().data_ptr(), C.data_ptr(), M, N, K,
                               torch.cuda.current_stream().cuda_stream)
    _ext.check(rc, "ggq_mm_ref_q8_1")
    return C


def mmq_q8_0_q8_1(A, B, M, N, K):
    return _call(_ext.GGQ_Q8_0, A, B, M, N, K)


def mmq_q4_k_q8_1(A, B, M, N, K):
    return _call(_ext.GGQ_Q4_K, A, B, M, N, K)


def mmq_q6_k_q8_1(A, B, M, N, K):
    return _call(_ext.GGQ_Q6_K, A, B, M, N, K)
