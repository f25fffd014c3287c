"""Loader + bindings of libggq.so (the C ABI in include/ggq.h).

Two bindings of the same C ABI: the per-step calls (`mm`, `mm_swiglu`) go through the PyTorch extension
`_ggq_torch.so` (csrc/torch_binding.cpp: operand checks, result allocation, current stream, ~2 us of host time per
call); everything else (queries, packers, dequantizers, the multi-GPU and host-pipe calls) through ctypes.
There is NO fallback: if a library is missing or a call fails, the op raises.  torch is used only for device
memory, the current stream and the device guard — the arithmetic is all in libggq.so.
"""
from __future__ import annotations

import ctypes
import os

import torch

_PKG = os.environ.get("GGQ_LIB_DIR") or os.path.dirname(os.path.dirname(os.path.abspath(__file__)))  # (dev: A/B builds)
_LIB_PATH = os.path.join(_PKG, "libggq.so")
_lib = None
_text = None

GGQ_Q8_0, GGQ_Q4_K, GGQ_Q6_K = 0, 1, 2
FAMILY_AUTO, FAMILY_GENERIC, FAMILY_DECODE, FAMILY_PREFILL, FAMILY_SKINNY = 0, 1, 2, 3, 4
FMT_ID = {"q8_0": GGQ_Q8_0, "q4_k": GGQ_Q4_K, "q6_k": GGQ_Q6_K}
FMT_QK = {GGQ_Q8_0: 32, GGQ_Q4_K: 256, GGQ_Q6_K: 256}
FMT_BLK = {GGQ_Q8_0: 34, GGQ_Q4_K: 144, GGQ_Q6_K: 210}

_I64, _P, _INT = ctypes.c_int64, ctypes.c_void_p, ctypes.c_int


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise RuntimeError(
                f"{_LIB_PATH} not found — the CUDA library is not built. Run `make -C {_PKG}` "
                "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback.")
        L = ctypes.CDLL(_LIB_PATH)
        for name in ("ggq_mm_q8_0_f16", "ggq_mm_q4_k_f16", "ggq_mm_q6_k_f16"):
            fn = getattr(L, name)
            fn.argtypes = [_P, _P, _P, _I64, _I64, _I64, _P]
            fn.restype = _INT
        L.ggq_mm_ex.argtypes = [_INT, _P, _P, _I64, ctypes.POINTER(_P), _INT, _I64, _I64, _I64, _I64, _INT, _P]
        L.ggq_mm_ex.restype = _INT
        for name in ("ggq_dequant_q8_0_f16", "ggq_dequant_q4_k_f16", "ggq_dequant_q6_k_f16"):
            fn = getattr(L, name)
            fn.argtypes = [_P, _P, _I64, _I64, _P]
            fn.restype = _INT
        L.ggq_packed_nbytes.argtypes = [_INT, _I64, _I64]
        L.ggq_packed_nbytes.restype = _I64
        L.ggq_select_family.argtypes = [_INT, _I64, _I64, _I64]
        L.ggq_select_family.restype = _INT
        L.ggq_launch_count.argtypes = []
        L.ggq_launch_count.restype = _I64
        L.ggq_error_string.argtypes = [_INT]
        L.ggq_error_string.restype = ctypes.c_char_p
        L.ggq_version.argtypes = []
        L.ggq_version.restype = _INT
        _lib = L
    return _lib


def torch_ext():
    """The PyTorch extension module (`_ggq_torch`), loaded from the package directory; raises when it is not built."""
    global _text
    if _text is None:
        lib()   # libggq.so first: the extension links it by name
        path = os.path.join(_PKG, "_ggq_torch.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} not found — the PyTorch extension is not built. Run `make -C {_PKG}`. "
                               "There is no fallback binding for the per-step calls.")
        import importlib.util
        spec = importlib.util.spec_from_file_location("_ggq_torch", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        if mod.version() != lib().ggq_version():
            raise RuntimeError("_ggq_torch.so and libggq.so were built from different sources; rebuild")
        _text = mod
    return _text


def check(rc: int, what: str) -> None:
    if rc != 0:
        raise RuntimeError(f"{what} failed ({rc}): {lib().ggq_error_string(rc).decode()}")


def _check_operands(fmt: int, A: torch.Tensor, B: torch.Tensor, M: int, N: int, K: int) -> None:
    if A.dtype not in (torch.int8, torch.uint8):
        raise TypeError(f"A must be int8 packed blocks, got {A.dtype}")
    if B.dtype != torch.float16:
        raise TypeError(f"B must be float16, got {B.dtype}")
    if not A.is_cuda or not B.is_cuda or A.device != B.device:
        raise ValueError("A and B must live on the same CUDA device (no CPU path)")
    if not A.is_contiguous() or not B.is_contiguous():
        raise ValueError("A and B must be contiguous")
    want = M * (K // FMT_QK[fmt]) * FMT_BLK[fmt]
    if A.numel() != want:
        raise ValueError(f"A has {A.numel()} bytes, expected {want} for M={M}, K={K}")
    if B.numel() != N * K:
        raise ValueError(f"B has {B.numel()} elements, expected N*K={N * K}")


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)  # cudaStream_t of a device's current stream


def _current_stream(index: int) -> int:
    if _raw_stream is not None:
        return _raw_stream(index)
    return torch.cuda.current_stream(index).cuda_stream


def mm(fmt: int, A: torch.Tensor, B: torch.Tensor, M: int, N: int, K: int, *, family: int = FAMILY_AUTO,
       out: torch.Tensor | None = None) -> torch.Tensor:
    """C[N, M] (fp16) = B[N, K] @ dequant(A)[M, K]^T on A's device, current stream, asynchronous."""
    return torch_ext().mm(fmt, A, B, M, N, K, family, out)


def mm_ctypes(fmt: int, A: torch.Tensor, B: torch.Tensor, M: int, N: int, K: int, *, family: int = FAMILY_AUTO,
              out: torch.Tensor | None = None) -> torch.Tensor:
    """The same call through the ctypes binding (tests: both bindings reach the same C entry point)."""
    _check_operands(fmt, A, B, M, N, K)
    C = torch.empty((N, M), device=A.device, dtype=torch.float16) if out is None else out
    if out is not None and (out.dtype != torch.float16 or out.shape != (N, M) or not out.is_contiguous()
                            or out.device != A.device):
        raise ValueError("out must be a contiguous float16 [N, M] tensor on A's device")
    outs = (_P * 1)(C.data_ptr())
    with torch.cuda.device(A.device):
        rc = lib().ggq_mm_ex(fmt, A.data_ptr(), B.data_ptr(), K, outs, 1, M, M, N, K, family, _current_stream(A.device.index))
    check(rc, "ggq_mm")
    return C


def mm_swiglu(fmt: int, Ag: torch.Tensor, Au: torch.Tensor, B: torch.Tensor, M: int, N: int, K: int) -> torch.Tensor:
    """C[N, M] (fp16) = silu(B @ dequant(Ag)^T) * (B @ dequant(Au)^T)  (ggq_mm_swiglu, include/ggq.h)."""
    return torch_ext().mm_swiglu(fmt, Ag, Au, B, M, N, K)


def mm_ex(fmt: int, A: torch.Tensor, B: torch.Tensor, outs: list[int], ldc: int, M: int, N: int, K: int, *,
          family: int = FAMILY_AUTO, ldx: int | None = None) -> None:
    """Raw form: `outs` are device pointers (own and peer-mapped) of fp16 [N, ldc] buffers."""
    arr = (_P * len(outs))(*outs)
    with torch.cuda.device(A.device):
        stream = torch.cuda.current_stream().cuda_stream
        rc = lib().ggq_mm_ex(fmt, A.data_ptr(), B.data_ptr(), K if ldx is None else ldx, arr, len(outs), ldc, M, N, K,
                             family, stream)
    check(rc, "ggq_mm_ex")


class PeerSync(ctypes.Structure):
    """ctypes mirror of `ggq_peer_sync` (include/ggq.h)."""
    _fields_ = [("rank", ctypes.c_int32), ("world", ctypes.c_int32), ("x_owner", ctypes.c_int32), ("epoch", ctypes.c_uint32),
                ("counter", _P), ("epoch_dev", _P), ("X_alt", _P), ("C_alt", _P),
                ("x_land", _P), ("x_land_peer", _P * 8), ("x_land_half", _I64),
                ("c_land", _P), ("c_land_peer", _P * 8), ("c_land_half", _I64),
                ("status", _P), ("timeout_ns", ctypes.c_uint64)]


def bind_mm_sync(L: ctypes.CDLL) -> None:
    L.ggq_mm_sync.argtypes = [_INT, _P, _P, _I64, _P, _I64, _I64, _I64, _I64, ctypes.POINTER(PeerSync),
                              ctypes.POINTER(_INT), _P]
    L.ggq_mm_sync.restype = _INT


def dequant(fmt: int, A: torch.Tensor, M: int, K: int) -> torch.Tensor:
    """fp16 [M, K] dequantized weights (bit-exact with the reference dequantizers)."""
    if A.dtype not in (torch.int8, torch.uint8) or not A.is_cuda or not A.is_contiguous():
        raise ValueError("A must be a contiguous int8 CUDA tensor")
    if A.numel() != M * (K // FMT_QK[fmt]) * FMT_BLK[fmt]:
        raise ValueError("packed size does not match M, K")
    out = torch.empty((M, K), device=A.device, dtype=torch.float16)
    with torch.cuda.device(A.device):
        stream = torch.cuda.current_stream().cuda_stream
        fn = (lib().ggq_dequant_q8_0_f16, lib().ggq_dequant_q4_k_f16, lib().ggq_dequant_q6_k_f16)[fmt]
        rc = fn(A.data_ptr(), out.data_ptr(), M, K, stream)
    check(rc, "ggq_dequant")
    return out


def launch_count() -> int:
    return int(lib().ggq_launch_count())
