"""GGUF (v2 / v3) container reader feeding the mmq ops — SURVEY §8(f) rank 3, the step before the path.

The reference only multiplies in-memory synthetic tensors (utils/quantize/q4_k.py:146-158, test/test_mmq_*.py);
real weights come in GGUF files.  This module parses the container (written from the published format
description, not from gguf-py) and hands Q8_0 / Q4_K / Q6_K tensors to the ops unchanged: a GGUF tensor of logical
shape [O, K] is stored row-major with K fastest, every row K/QK blocks — exactly the flat byte stream the reference
packers emit and `mmq_*(A, B, M, N, K)` consumes, so loading is an mmap slice and one H2D copy, no repacking.

File layout (little endian):
    "GGUF" | u32 version | u64 n_tensors | u64 n_kv
    n_kv   x { string key | u32 value_type | value }
    n_tens x { string name | u32 n_dims | u64 dims[n_dims] (fastest first) | u32 ggml_type | u64 offset }
    padding to `general.alignment` (default 32)
    tensor data, each tensor at data_start + offset
string = u64 length + UTF-8 bytes;  array value = u32 element_type | u64 count | elements.
"""
from __future__ import annotations

import mmap
import struct
from dataclasses import dataclass
from typing import Any

import numpy as np

GGUF_MAGIC = b"GGUF"
DEFAULT_ALIGNMENT = 32

# metadata value types
_SCALARS = {0: "<B", 1: "<b", 2: "<H", 3: "<h", 4: "<I", 5: "<i", 6: "<f", 7: "<?", 10: "<Q", 11: "<q", 12: "<d"}
_T_STRING, _T_ARRAY = 8, 9

# ggml tensor types: id -> (name, elements per block, bytes per block).  Q8_1: gguf-py's table (the file-format authority
# this reader is tested against) says 40 B (fp32 d and s); ggml's in-memory block_q8_1 — what ggq_quantize_q8_1_f16 and the
# reference's q8_1.py produce — is 36 B (fp16 d, s).  Q8_1 is an activation format and does not occur in weight files.
GGML_TYPES = {
    0: ("F32", 1, 4), 1: ("F16", 1, 2), 2: ("Q4_0", 32, 18), 3: ("Q4_1", 32, 20), 6: ("Q5_0", 32, 22),
    7: ("Q5_1", 32, 24), 8: ("Q8_0", 32, 34), 9: ("Q8_1", 32, 40), 10: ("Q2_K", 256, 84), 11: ("Q3_K", 256, 110),
    12: ("Q4_K", 256, 144), 13: ("Q5_K", 256, 176), 14: ("Q6_K", 256, 210), 15: ("Q8_K", 256, 292),
    24: ("I8", 1, 1), 25: ("I16", 1, 2), 26: ("I32", 1, 4), 27: ("I64", 1, 8), 28: ("F64", 1, 8), 30: ("BF16", 1, 2),
}
# the formats the mmq path multiplies: ggml type id -> ggq format name
MMQ_FORMATS = {8: "q8_0", 12: "q4_k", 14: "q6_k"}


class GGUFError(ValueError):
    pass


@dataclass(frozen=True)
class TensorInfo:
    name: str
    shape: tuple[int, ...]      # logical shape, slowest dimension first (a weight is [O, K])
    ggml_type: int
    type_name: str
    offset: int                 # absolute byte offset in the file
    nbytes: int

    @property
    def mmq_format(self) -> str | None:
        return MMQ_FORMATS.get(self.ggml_type)


class _Cursor:
    def __init__(self, buf, pos: int = 0):
        self.buf, self.pos = buf, pos

    def take(self, fmt: str):
        size = struct.calcsize(fmt)
        if self.pos + size > len(self.buf):
            raise GGUFError("truncated GGUF header")
        (v,) = struct.unpack_from(fmt, self.buf, self.pos)
        self.pos += size
        return v

    def string(self) -> str:
        n = self.take("<Q")
        if self.pos + n > len(self.buf):
            raise GGUFError("truncated GGUF string")
        s = bytes(self.buf[self.pos:self.pos + n]).decode("utf-8")
        self.pos += n
        return s

    def value(self, vtype: int) -> Any:
        if vtype in _SCALARS:
            return self.take(_SCALARS[vtype])
        if vtype == _T_STRING:
            return self.string()
        if vtype == _T_ARRAY:
            etype = self.take("<I")
            count = self.take("<Q")
            if etype in _SCALARS and etype != 7:   # numeric arrays in one go
                fmt = _SCALARS[etype][1]
                size = struct.calcsize("<" + fmt) * count
                if self.pos + size > len(self.buf):
                    raise GGUFError("truncated GGUF array")
                arr = np.frombuffer(self.buf, dtype=np.dtype("<" + fmt), count=count, offset=self.pos).tolist()
                self.pos += size
                return arr
            return [self.value(etype) for _ in range(count)]
        raise GGUFError(f"unknown GGUF metadata value type {vtype}")


class GGUFFile:
    """Memory-mapped GGUF file: `.metadata` (dict), `.tensors` (name -> TensorInfo), zero-copy tensor bytes."""

    def __init__(self, path: str):
        self.path = path
        self._f = open(path, "rb")
        try:
            self._mm = mmap.mmap(self._f.fileno(), 0, access=mmap.ACCESS_READ)
        except ValueError as e:          # empty file
            self._f.close()
            raise GGUFError(f"{path}: not a GGUF file ({e})") from None
        c = _Cursor(self._mm)
        if len(self._mm) < 24 or bytes(self._mm[:4]) != GGUF_MAGIC:
            self.close()
            raise GGUFError(f"{path}: bad magic, not a GGUF file")
        c.pos = 4
        self.version = c.take("<I")
        if self.version not in (2, 3):
            self.close()
            raise GGUFError(f"{path}: unsupported GGUF version {self.version} (2 and 3 are supported)")
        n_tensors, n_kv = c.take("<Q"), c.take("<Q")
        self.metadata: dict[str, Any] = {}
        for _ in range(n_kv):
            key = c.string()
            self.metadata[key] = c.value(c.take("<I"))
        raw = []
        for _ in range(n_tensors):
            name = c.string()
            n_dims = c.take("<I")
            dims = [c.take("<Q") for _ in range(n_dims)]
            raw.append((name, dims, c.take("<I"), c.take("<Q")))
        self.alignment = int(self.metadata.get("general.alignment", DEFAULT_ALIGNMENT))
        if self.alignment <= 0 or self.alignment & (self.alignment - 1):
            self.close()
            raise GGUFError(f"{path}: general.alignment = {self.alignment} is not a power of two")
        self.data_start = (c.pos + self.alignment - 1) // self.alignment * self.alignment
        self.tensors: dict[str, TensorInfo] = {}
        for name, dims, gtype, rel in raw:
            if gtype not in GGML_TYPES:
                tname, nbytes = f"type{gtype}", -1          # unknown type: listed, not loadable
            else:
                tname, qk, blk = GGML_TYPES[gtype]
                n = int(np.prod(dims, dtype=np.int64)) if dims else 1
                if dims and dims[0] % qk:
                    self.close()
                    raise GGUFError(f"{path}: tensor {name}: row length {dims[0]} is not a multiple of {qk} ({tname})")
                nbytes = n // qk * blk
            off = self.data_start + rel
            if rel % self.alignment or (nbytes >= 0 and off + nbytes > len(self._mm)):
                self.close()
                raise GGUFError(f"{path}: tensor {name}: data [{off}, {off + max(nbytes, 0)}) is misaligned or past the end of the file")
            self.tensors[name] = TensorInfo(name, tuple(reversed(dims)), gtype, tname, off, nbytes)

    # ---- raw access -------------------------------------------------------------------------------------
    def tensor_bytes(self, name: str) -> np.ndarray:
        """Zero-copy uint8 view of the tensor's bytes in the mapped file."""
        t = self.tensors[name]
        if t.nbytes < 0:
            raise GGUFError(f"tensor {name}: ggml type {t.ggml_type} is not known to this reader")
        return np.frombuffer(self._mm, dtype=np.uint8, count=t.nbytes, offset=t.offset)

    def tensor_numpy(self, name: str) -> np.ndarray:
        """F32 / F16 / F64 / integer tensors as a numpy array of their logical shape (zero-copy)."""
        t = self.tensors[name]
        dt = {"F32": "<f4", "F16": "<f2", "F64": "<f8", "I8": "i1", "I16": "<i2", "I32": "<i4", "I64": "<i8"}.get(t.type_name)
        if dt is None:
            raise GGUFError(f"tensor {name} is {t.type_name}: use tensor_bytes() / load_quantized()")
        return self.tensor_bytes(name).view(dt).reshape(t.shape)

    # ---- feeding the ops --------------------------------------------------------------------------------
    def load_quantized(self, name: str, device="cuda"):
        """(fmt, A, O, K): the packed int8 tensor `mmq_<fmt>(A, B, O, T, K)` takes, on `device`."""
        import torch

        t = self.tensors[name]
        fmt = t.mmq_format
        if fmt is None:
            raise GGUFError(f"tensor {name} is {t.type_name}; the mmq path multiplies Q8_0, Q4_K and Q6_K")
        if len(t.shape) != 2:
            raise GGUFError(f"tensor {name} has shape {t.shape}; a weight matrix [O, K] is expected")
        O, K = t.shape
        host = torch.from_numpy(self.tensor_bytes(name).view(np.int8).copy())   # the mmap is read-only: one host copy
        return fmt, host.to(device), int(O), int(K)

    def linear(self, name: str, device="cuda"):
        """A callable X[T, K] -> X @ W^T [T, O] over the named quantized tensor (reference-named entry points)."""
        from kernels.mmq_q4_k import mmq_q4_k
        from kernels.mmq_q6_k import mmq_q6_k
        from kernels.mmq_q8_0 import mmq_q8_0

        fmt, A, O, K = self.load_quantized(name, device)
        fn = {"q8_0": mmq_q8_0, "q4_k": mmq_q4_k, "q6_k": mmq_q6_k}[fmt]

        def apply(X):
            if X.dim() != 2 or X.shape[1] != K:
                raise ValueError(f"X must be [T, {K}], got {tuple(X.shape)}")
            return fn(A, X.contiguous(), O, X.shape[0], K)

        apply.fmt, apply.O, apply.K, apply.weight = fmt, O, K, A
        return apply

    def close(self) -> None:
        mm, f = getattr(self, "_mm", None), getattr(self, "_f", None)
        self._mm = None
        if mm is not None:
            try:
                mm.close()
            except BufferError:   # numpy views of the mapping are still alive; the OS unmaps at exit
                pass
        if f is not None:
            f.close()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
