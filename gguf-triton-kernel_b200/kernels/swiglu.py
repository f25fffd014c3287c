"""Fused SwiGLU up-projection over GGUF-packed gate / up weights.

No reference counterpart: a caller of the reference runs ``mmq_*`` twice and applies ``silu(gate) * up`` in torch.
``mmq_swiglu`` returns exactly that value — ``(F.silu(mmq(Ag, B).float()) * mmq(Au, B).float()).half()`` — from one
kernel for decode-sized token counts (``ggq_mm_swiglu``, include/ggq.h): both packed matrices are streamed once and
the activation is applied in the accumulator registers.
"""
import torch

from . import _ext


def mmq_swiglu(fmt: str, A_gate: torch.Tensor, A_up: torch.Tensor, B: torch.Tensor, M: int, N: int, K: int) -> torch.Tensor:
    """
    Args:
        fmt: "q8_0" | "q4_k" | "q6_k" (both weights use it)
        A_gate, A_up: packed weights, flat int8, ``[M, K]`` logical each, on a CUDA device
        B: fp16 ``[N, K]`` on the same device
    Returns:
        fp16 ``[N, M]`` = silu(B @ dequant(A_gate).T) * (B @ dequant(A_up).T)
    """
    f = _ext.FMT_ID[fmt]
    assert (K % _ext.FMT_QK[f] == 0)
    return _ext.mm_swiglu(f, A_gate, A_up, B, M, N, K)


def mmq_q8_0_swiglu(A_gate, A_up, B, M, N, K):
    return mmq_swiglu("q8_0", A_gate, A_up, B, M, N, K)


def mmq_q4_k_swiglu(A_gate, A_up, B, M, N, K):
    return mmq_swiglu("q4_k", A_gate, A_up, B, M, N, K)


def mmq_q6_k_swiglu(A_gate, A_up, B, M, N, K):
    return mmq_swiglu("q6_k", A_gate, A_up, B, M, N, K)
