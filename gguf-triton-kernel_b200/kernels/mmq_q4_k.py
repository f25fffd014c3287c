"""Q4_K @ fp16 — drop-in for the reference's ``kernels/mmq_q4_k.py`` (``mmq_q4_k`` at :240-289).

Same name, positional signature, module constants (:232-236), operand layouts and result layout; the
body calls the C ABI (``ggq_mm_q4_k_f16``, include/ggq.h) over hand-written sm_100a CUDA.
"""
import torch

from . import _ext

Q4_K_BLOCK_SIZE = 144  # bytes
Q8_1_BLOCK_SIZE = 36  # bytes
Q4_K_SUBBLK_NUM = 8
QK_K = 256
QK8_1 = 32


def mmq_q4_k(A: torch.Tensor, B: torch.Tensor, M: int, N: int, K: int) -> torch.Tensor:
    """out = (A @ B.T).T

    Args:
        A: Q4_K packed weight, flat int8 ``[M * K/256 * 144]`` on a CUDA device
        B: fp16 ``[N, K]`` on the same device
        M: rows of A (out-features);  N: rows of B (tokens);  K: columns of both
    Returns:
        fp16 ``[N, M]``, contiguous, on ``A.device``
    """
    assert (K % 256 == 0)
    return _ext.mm(_ext.GGQ_Q4_K, A, B, M, N, K)
