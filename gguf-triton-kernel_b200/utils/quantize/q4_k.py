"""GPU mirrors of the reference's utils/quantize/q4_k.py: dequantizer (bit-exact fp16) and packer (byte-identical)."""
import torch

from ._common import dequant, quantize_k


def dequantize_q4_k(quantized_tensor: torch.Tensor, original_shape) -> torch.Tensor:
    return dequant("q4_k", quantized_tensor, original_shape, 144, 256)


def quantize_to_q4_k(input_tensor: torch.Tensor) -> torch.Tensor:
    """GPU packer, byte-identical to the reference's compiled `quantize_row_q4_K_ref` behind
    utils/quantize/q4_k.py:86-90: flat int8 [n/256 * 144] on the input's device."""
    return quantize_k("ggq_quantize_q4_k_f32", 144, input_tensor)
