import ctypes

import torch

from kernels import _ext

_P, _I64 = ctypes.c_void_p, ctypes.c_int64
_ready = False


def _lib():
    global _ready
    L = _ext.lib()
    if not _ready:
        for name in ("ggq_quantize_q8_0_f16", "ggq_quantize_q8_1_f16", "ggq_quantize_q4_k_f32", "ggq_quantize_q6_k_f32"):
            fn = getattr(L, name)
            fn.argtypes = [_P, _P, _I64, _P]
            fn.restype = ctypes.c_int
        L.ggq_dequant_q6_k_f32.argtypes = [_P, _P, _I64, _I64, _P]
        L.ggq_dequant_q6_k_f32.restype = ctypes.c_int
        _ready = True
    return L


def quantize_q8(name: str, blk: int, x: torch.Tensor) -> torch.Tensor:
    if not x.is_cuda:
        raise ValueError("CUDA tensor expected (no CPU path)")
    if x.dtype != torch.float16:
        x = x.to(torch.float16)
    flat = x.contiguous().flatten()
    n = flat.numel()
    if n % 32 != 0:
        raise ValueError("The total number of elements must be divisible by 32.")
    out = torch.empty(n // 32 * blk, dtype=torch.int8, device=x.device)
    with torch.cuda.device(x.device):
        rc = getattr(_lib(), name)(flat.data_ptr(), out.data_ptr(), n, torch.cuda.current_stream().cuda_stream)
    _ext.check(rc, name)
    return out


def quantize_k(name: str, blk: int, x: torch.Tensor) -> torch.Tensor:
    """K-quant packers: the reference converts its input to float32 and packs 256-element super-blocks
    (utils/quantize/q4_k.py:86-90, q6_k.py:99-110); same here, on the GPU, byte-identical."""
    if not x.is_cuda:
        raise ValueError("CUDA tensor expected (no CPU path)")
    flat = x.to(torch.float32).contiguous().flatten()
    n = flat.numel()
    if n % 256 != 0:
        raise ValueError(f"Array length must be multiple of 256 (got {n})")
    out = torch.empty(n // 256 * blk, dtype=torch.int8, device=x.device)
    with torch.cuda.device(x.device):
        rc = getattr(_lib(), name)(flat.data_ptr(), out.data_ptr(), n, torch.cuda.current_stream().cuda_stream)
    _ext.check(rc, name)
    return out


def dequant(fmt: str, q: torch.Tensor, shape, blk: int, qk: int) -> torch.Tensor:
    if q.dtype != torch.int8:
        raise ValueError("Quantized tensor must be of type int8")
    if not q.is_cuda:
        raise ValueError("CUDA tensor expected (no CPU path)")
    n = q.numel()
    if n % blk != 0:
        raise ValueError(f"Invalid quantized tensor size. Expected size divisible by {blk}, got {n}.")
    nb = n // blk
    # rows of at most 2^20 blocks: the C ABI carries K as a 32-bit quantity inside its kernels (GGQ_E_SHAPE beyond)
    per = nb
    while per > (1 << 20) and per % 2 == 0:
        per //= 2
    return _ext.dequant(_ext.FMT_ID[fmt], q.contiguous(), nb // per, per * qk).reshape(shape)
