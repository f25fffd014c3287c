"""K-quant packers for synthetic test weights.  TEST INFRASTRUCTURE ONLY (see ggq_oracle.py).

``quantize_to_q4_k`` / ``quantize_to_q6_k`` call the reference's own C packers
(``quantize_row_q4_K_ref`` utils/quantize/q4_k_ref.c:281-368, ``quantize_row_q6_K_ref``
utils/quantize/q6_k_ref.c:243-340) compiled by ``oracle/Makefile`` from ``/root/reference`` into
``oracle/_ref/`` — the same call the reference wrappers make (utils/quantize/q4_k.py:60-91,
q6_k.py:70-108): fp16 tensor -> fp32 -> one C call over the flattened array -> raw block bytes.
Rows are independent (K % 256 == 0), so large tensors are packed by a thread pool over row chunks
(ctypes releases the GIL); the bytes are identical to a single call.
"""
from __future__ import annotations

import ctypes
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from .ggq_oracle import Q4_K_SIZE, Q6_K_SIZE, QK_K, quantize_to_q8_0  # noqa: F401  (re-export)

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIBS: dict[str, ctypes.CDLL] = {}


def ref_packers_available() -> bool:
    return all(os.path.exists(os.path.join(_HERE, "_ref", n))
               for n in ("libq4_k_ref.so", "libq6_k_ref.so"))


def _lib(name: str, sym: str):
    if name not in _LIBS:
        path = os.path.join(_HERE, "_ref", name)
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing: run `make -C oracle ref` where /root/reference exists")
        lib = ctypes.CDLL(path)
        fn = getattr(lib, sym)
        fn.argtypes = [ctypes.POINTER(ctypes.c_float), ctypes.c_void_p, ctypes.c_int64]
        fn.restype = None
        _LIBS[name] = lib
    return getattr(_LIBS[name], sym)


def _pack(x, libname: str, sym: str, blk: int, threads: int | None) -> np.ndarray:
    arr = np.ascontiguousarray(np.asarray(x).astype(np.float16).astype(np.float32)).reshape(-1)
    n = arr.size
    if n % QK_K:
        raise ValueError(f"element count must be a multiple of {QK_K} (got {n})")
    nblk = n // QK_K
    out = np.empty(nblk * blk, dtype=np.uint8)
    fn = _lib(libname, sym)

    def run(b0: int, b1: int):
        src = arr[b0 * QK_K:b1 * QK_K]
        dst = out[b0 * blk:b1 * blk]
        fn(src.ctypes.data_as(ctypes.POINTER(ctypes.c_float)), dst.ctypes.data, ctypes.c_int64(src.size))

    threads = threads or min(32, os.cpu_count() or 1)
    if nblk < 4096 or threads <= 1:
        run(0, nblk)
    else:
        step = -(-nblk // (threads * 4))
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(lambda b0: run(b0, min(nblk, b0 + step)), range(0, nblk, step)))
    return out.view(np.int8)


def quantize_to_q4_k(x, threads: int | None = None) -> np.ndarray:
    return _pack(x, "libq4_k_ref.so", "quantize_row_q4_K_ref", Q4_K_SIZE, threads)


def quantize_to_q6_k(x, threads: int | None = None) -> np.ndarray:
    return _pack(x, "libq6_k_ref.so", "quantize_row_q6_K_ref", Q6_K_SIZE, threads)


def quantize(fmt: str, x, threads: int | None = None) -> np.ndarray:
    if fmt == "q8_0":
        return quantize_to_q8_0(x)
    if fmt == "q4_k":
        return quantize_to_q4_k(x, threads)
    if fmt == "q6_k":
        return quantize_to_q6_k(x, threads)
    raise KeyError(fmt)
