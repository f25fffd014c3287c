"""GPU-backed mirrors of the reference's ``utils/quantize`` functions that sit on either side of the mmq path
(SURVEY §8f "next" rows): same names, same return dtypes/shapes, CUDA tensors in and out, no CPU fallback.

    quantize_to_q8_0 / dequantize_q8_0      utils/quantize/q8_0.py:4-100
    quantize_to_q8_1                        utils/quantize/q8_1.py:18-70
    quantize_to_q4_k / dequantize_q4_k      utils/quantize/q4_k.py:86-90, 146-158 (+ q4_k_ref.c:188-368)
    quantize_to_q6_k / dequantize_q6_k      utils/quantize/q6_k.py:99-110, 138-159 (+ q6_k_ref.c:153-340; fp32 dequant)

The packers are byte-identical to the reference's (the K-quant ones reproduce GGML's iterative scale search one
thread per sub-block, csrc/kquant_pack.cuh).
"""
