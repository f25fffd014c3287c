"""B200-native GGUF mmq kernels behind the reference's Python entry points.

    from kernels.mmq_q8_0 import mmq_q8_0
    from kernels.mmq_q4_k import mmq_q4_k
    from kernels.mmq_q6_k import mmq_q6_k
"""
