"""Mirror of the reference's utils/quantize/q8_0.py on the GPU (byte-identical packer, bit-exact dequantizer)."""
import torch

from ._common import dequant, quantize_q8


def quantize_to_q8_0(input_tensor: torch.Tensor) -> torch.Tensor:
    return quantize_q8("ggq_quantize_q8_0_f16", 34, input_tensor)


def dequantize_q8_0(quantized_tensor: torch.Tensor, original_shape) -> torch.Tensor:
    return dequant("q8_0", quantized_tensor, original_shape, 34, 32)
