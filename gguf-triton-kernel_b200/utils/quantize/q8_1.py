"""Mirror of the reference's utils/quantize/q8_1.py packer on the GPU (byte-identical)."""
import torch

from ._common import quantize_q8


def quantize_to_q8_1(input_tensor: torch.Tensor) -> torch.Tensor:
    return quantize_q8("ggq_quantize_q8_1_f16", 36, input_tensor)
