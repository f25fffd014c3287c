"""GPU mirror of the reference's utils/quantize/q6_k.py dequantizer: returns float32 like the reference (:157)."""
import torch

from kernels import _ext
from ._common import _lib


def dequantize_q6_k(quantized_tensor: torch.Tensor, original_shape) -> torch.Tensor:
    q = quantized_tensor
    if not q.is_cuda:
        raise ValueError("CUDA tensor expected (no CPU path)")
    n = q.numel()
    if n % 210 != 0:
        raise ValueError(f"Invalid quantized tensor size. Expected size divisible by 210, got {n}.")
    q = q.contiguous()
    out = torch.empty(n // 210 * 256, dtype=torch.float32, device=q.device)
    with torch.cuda.device(q.device):
        rc = _lib().ggq_dequant_q6_k_f32(q.data_ptr(), out.data_ptr(), 1, out.numel(), torch.cuda.current_stream().cuda_stream)
    _ext.check(rc, "ggq_dequant_q6_k_f32")
    return out.reshape(original_shape)
