"""Q6_K @ fp16 — drop-in for the reference's ``kernels/mmq_q6_k.py`` (``mmq_q6_k`` at :197-246).

Same name, positional signature, module constants (:189-193), operand layouts and result layout; the
body calls the C ABI (``ggq_mm_q6_k_f16``, include/ggq.h) over hand-written sm_100a CUDA.
"""
import torch

from . import _ext

QK_K = 256
Q6_K_SUBBLK_NUM = 16
QK8_1 = 32
Q6_K_BLOCK_SIZE = 210  # bytes
Q8_1_BLOCK_SIZE = 36  # bytes


def mmq_q6_k(A: torch.Tensor, B: torch.Tensor, M: int, N: int, K: int) -> torch.Tensor:
    """out = (A @ B.T).T

    Args:
        A: Q6_K packed weight, flat int8 ``[M * K/256 * 210]`` on a CUDA device
        B: fp16 ``[N, K]`` on the same device
        M: rows of A (out-features);  N: rows of B (tokens);  K: columns of both
    Returns:
        fp16 ``[N, M]``, contiguous, on ``A.device``
    """
    assert (K % 256 == 0)
    return _ext.mm(_ext.GGQ_Q6_K, A, B, M, N, K)
