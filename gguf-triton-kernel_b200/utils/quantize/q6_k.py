"""GPU mirrors of the reference's utils/quantize/q6_k.py: dequantizer (float32 like the reference, :157) and packer."""
import torch

from kernels import _ext
from ._common import _lib, quantize_k


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


def quantize_to_q6_k(input_tensor: torch.Tensor) -> torch.Tensor:
    """GPU packer, byte-identical to the reference's compiled `quantize_row_q6_K_ref` behind
    utils/quantize/q6_k.py:99-110: flat int8 [n/256 * 210] on the input's device."""
    return quantize_k("ggq_quantize_q6_k_f32", 210, input_tensor)
