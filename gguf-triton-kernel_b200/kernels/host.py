"""The mmq step with HOST activations and a HOST result (``ggq_host_pipe`` / ``ggq_mm_host``, include/ggq.h).

What a caller of the reference does around ``mmq_*`` — ``B.cuda()`` before, ``.cpu()`` after — behind one call: the
pipe owns rotating device slots and three streams, so the copy-in of the next step and the copy-out of the previous
one overlap the kernel of the current step.  Everything is enqueued asynchronously; ``sync()`` waits.
"""
from __future__ import annotations

import ctypes

import torch

from . import _ext

_I64, _P, _INT = ctypes.c_int64, ctypes.c_void_p, ctypes.c_int
_bound = False


def _lib():
    global _bound
    L = _ext.lib()
    if not _bound:
        L.ggq_host_pipe_create.argtypes = [ctypes.POINTER(_P), _I64, _I64, _INT]
        L.ggq_host_pipe_create.restype = _INT
        L.ggq_mm_host.argtypes = [_P, _INT, _P, _P, _P, _I64, _I64, _I64]
        L.ggq_mm_host.restype = _INT
        L.ggq_host_pipe_sync.argtypes = [_P]
        L.ggq_host_pipe_sync.restype = _INT
        L.ggq_host_pipe_stream.argtypes = [_P, _INT]
        L.ggq_host_pipe_stream.restype = _P
        L.ggq_host_pipe_destroy.argtypes = [_P]
        L.ggq_host_pipe_destroy.restype = None
        _bound = True
    return L


class HostPipe:
    """``pipe = HostPipe(fmt, A, M, K, max_tokens)``; ``pipe(B_host, out=C_host)`` per step; ``pipe.sync()``.

    A: packed weights on a CUDA device (the layer's resident state).  B_host: fp16 ``[N, K]`` CPU tensor (pinned memory
    makes the copies asynchronous); the result lands in ``out`` (fp16 ``[N, M]`` CPU tensor, allocated pinned when
    omitted) and is complete after ``sync()``.
    """

    def __init__(self, fmt: str, A: torch.Tensor, M: int, K: int, max_tokens: int, depth: int = 3):
        f = _ext.FMT_ID[fmt]
        assert (K % _ext.FMT_QK[f] == 0)
        if A.dtype not in (torch.int8, torch.uint8) or not A.is_cuda or not A.is_contiguous():
            raise ValueError("A must be a contiguous int8 CUDA tensor")
        if A.numel() != M * (K // _ext.FMT_QK[f]) * _ext.FMT_BLK[f]:
            raise ValueError("packed size does not match M, K")
        self.fmt, self.A, self.M, self.K, self.max_tokens = f, A, M, K, max_tokens
        self._h = _P()
        with torch.cuda.device(A.device):
            _ext.check(_lib().ggq_host_pipe_create(ctypes.byref(self._h), max(1, max_tokens * K * 2),
                                                   max(1, max_tokens * M * 2), depth), "ggq_host_pipe_create")

    def __call__(self, B_host: torch.Tensor, out: torch.Tensor | None = None) -> torch.Tensor:
        if B_host.is_cuda or B_host.dtype != torch.float16 or not B_host.is_contiguous() or B_host.dim() != 2 \
                or B_host.shape[1] != self.K:
            raise ValueError("B_host must be a contiguous float16 [N, K] CPU tensor")
        N = B_host.shape[0]
        if N > self.max_tokens:
            raise ValueError(f"{N} tokens, the pipe was created for {self.max_tokens}")
        if out is None:
            out = torch.empty((N, self.M), dtype=torch.float16).pin_memory()
        elif out.is_cuda or out.dtype != torch.float16 or not out.is_contiguous() or out.shape != (N, self.M):
            raise ValueError("out must be a contiguous float16 [N, M] CPU tensor")
        dev = self.A.device.index
        if torch.cuda.current_device() == dev:
            rc = _lib().ggq_mm_host(self._h, self.fmt, self.A.data_ptr(), B_host.data_ptr(), out.data_ptr(), self.M, N, self.K)
        else:
            with torch.cuda.device(dev):
                rc = _lib().ggq_mm_host(self._h, self.fmt, self.A.data_ptr(), B_host.data_ptr(), out.data_ptr(), self.M, N,
                                        self.K)
        _ext.check(rc, "ggq_mm_host")
        return out

    def sync(self) -> None:
        _ext.check(_lib().ggq_host_pipe_sync(self._h), "ggq_host_pipe_sync")

    def stream(self, which: int) -> torch.cuda.ExternalStream:
        """0 = copy-in, 1 = kernels, 2 = copy-out (to record events around a region)."""
        return torch.cuda.ExternalStream(_lib().ggq_host_pipe_stream(self._h, which), device=self.A.device)

    def close(self) -> None:
        if self._h:
            _lib().ggq_host_pipe_destroy(self._h)
            self._h = _P()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
