"""Q8_0 @ fp16 — drop-in for the reference's ``kernels/mmq_q8_0.py`` (``mmq_q8_0`` at :102-147).

Same name, positional signature, module constants (:96-98), operand layouts and result layout; the
body calls the C ABI (``ggq_mm_q8_0_f16``, include/ggq.h) over hand-written sm_100a CUDA instead of
launching a Triton kernel.  No Triton, no CPU fallback.
"""
import torch

from . import _ext

QK8_0 = 32  # weights per block
QK8_1 = 32
Q8_0_SIZE = 34  # bytes


def mmq_q8_0(A: torch.Tensor, B: torch.Tensor, M: int, N: int, K: int) -> torch.Tensor:
    """out = (A @ B.T).T

    Args:
        A: Q8_0 packed weight, flat int8 ``[M * K/32 * 34]`` on a CUDA device
        B: fp16 ``[N, K]`` on the same device
        M: rows of A (out-features);  N: rows of B (tokens);  K: columns of both
    Returns:
        fp16 ``[N, M]``, contiguous, on ``A.device`` (enqueued on the current stream, not synchronised)
    """
    assert (K % 32 == 0)
    return _ext.mm(_ext.GGQ_Q8_0, A, B, M, N, K)
