"""GPU-backed mirrors of the reference's ``utils/quantize`` functions that sit on either side of the mmq path
(SURVEY §8f "next" rows): same names, same return dtypes/shapes, CUDA tensors in and out, no CPU fallback.

    quantize_to_q8_0 / dequantize_q8_0      utils/quantize/q8_0.py:4-100
    quantize_to_q8_1                        utils/quantize/q8_1.py:18-70
    dequantize_q4_k                         utils/quantize/q4_k.py:146-158
    dequantize_q6_k (fp32, like the ref)    utils/quantize/q6_k.py:138-159

The K-quant *packers* (quantize_to_q4_k / q6_k, GGML's iterative search) are not provided here.
"""
