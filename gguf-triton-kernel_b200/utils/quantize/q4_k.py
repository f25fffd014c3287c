"""GPU mirror of the reference's utils/quantize/q4_k.py dequantizer (bit-exact fp16)."""
import torch

from ._common import dequant


def dequantize_q4_k(quantized_tensor: torch.Tensor, original_shape) -> torch.Tensor:
    return dequant("q4_k", quantized_tensor, original_shape, 144, 256)
