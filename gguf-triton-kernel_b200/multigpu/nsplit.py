"""N-split of one quantized layer across G GPUs of one node (one process per GPU).

The reference has no multi-GPU code; this is the BASELINE config-5 layout.  Rows of the packed weight
are independent units (no block straddles a row: K % QK == 0, kernels/mmq_q4_k.py:263), so rank r
simply owns the contiguous byte range of rows [r*O/G, (r+1)*O/G).  Activations are replicated
(broadcast from rank 0), every rank computes C_r[T, O/G] with the single-GPU kernels, and the slices
are exchanged so that every rank ends up with the full C[T, O].  Two exchange paths:

  "nccl"   ncclAllGather of the contiguous [T, O/G] slices into [G, T, O/G], then (T > 1) one transpose
           copy to [T, O].  Baseline path.
  "fused"  no collective kernel at all.
           T <= 8 (and T*K <= 65536): the whole step is ONE decode kernel per rank (ggq_mm_sync).  Rank 0's kernel
           pushes the activations to the peers and every rank sends its finished output tiles to all peers as
           flag-in-data lines (16-byte stores carrying 4 fp16 + the step's epoch) into peer-mapped landing buffers
           (torch symmetric memory supplies the NVLink mappings); the receivers poll the lines themselves, so each
           transfer is one NVLink hop with no fence and no flag round trip, and kernel completion == C[T, O] complete
           on this rank.  The epoch lives in device memory, so a step is CUDA-graph capturable.
           More tokens: NCCL broadcast of X, then between two symmetric-memory barriers either (slices under 1 MB) the
           GEMM's epilogue peer-stores each output tile into the [T, O] buffers of ALL ranks (n_out outputs, ldc = O),
           or (prefill-sized slices) the GEMM writes the rank's own buffer and the copy engines push the [T, O/N]
           column block into every peer's buffer (ggq_push_columns: one 2-D DMA copy per peer over NVLink).

`mm_fn` is injectable so the host logic (sharding arithmetic, exchange layout) is testable on CPU with
the gloo backend and a stand-in matmul.
"""
from __future__ import annotations

from typing import Callable

import torch
import torch.distributed as dist

FMT_QK = {"q8_0": 32, "q4_k": 256, "q6_k": 256}
FMT_BLK = {"q8_0": 34, "q4_k": 144, "q6_k": 210}


def row_bytes(fmt: str, K: int) -> int:
    if K % FMT_QK[fmt]:
        raise ValueError(f"K={K} is not a multiple of the {fmt} block size")
    return K // FMT_QK[fmt] * FMT_BLK[fmt]


def shard_rows(O: int, world: int, rank: int) -> tuple[int, int]:
    """Row range of `rank`.  O must divide evenly (true for every BASELINE shape: SURVEY §8e)."""
    if O % world:
        raise ValueError(f"O={O} is not divisible by the {world} ranks")
    per = O // world
    return rank * per, (rank + 1) * per


def shard_packed(fmt: str, A: torch.Tensor, O: int, K: int, world: int, rank: int) -> torch.Tensor:
    """The byte slice of the flat packed tensor `A` that holds this rank's rows (a view, no copy)."""
    rb = row_bytes(fmt, K)
    if A.numel() != O * rb:
        raise ValueError("packed size does not match O, K")
    lo, hi = shard_rows(O, world, rank)
    return A[lo * rb:hi * rb]


def assemble(gathered: torch.Tensor) -> torch.Tensor:
    """[G, T, O/G] (all-gather order) -> [T, O] contiguous."""
    G, T, per = gathered.shape
    if T == 1:
        return gathered.reshape(1, G * per)
    return gathered.permute(1, 0, 2).reshape(T, G * per)


class NSplitLinear:
    DECODE_MAX_T = 8   # fused decode kernel for T <= 8; 9..16 tokens take the GEMM path (tcgen05 skinny kernel + peer stores)

    def __init__(self, fmt: str, A_shard: torch.Tensor, O: int, K: int, *, group=None, mode: str = "nccl",
                 max_tokens: int = 16, mm_fn: Callable | None = None, sync_timeout_s: float = 2.0):
        self.fmt, self.O, self.K = fmt, O, K
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        self.lo, self.hi = shard_rows(O, self.world, self.rank)
        self.per = self.hi - self.lo
        if A_shard.numel() != self.per * row_bytes(fmt, K):
            raise ValueError("A_shard does not hold exactly this rank's rows")
        self.A = A_shard
        self.mode = mode
        self.max_tokens = max_tokens
        self.mm_fn = mm_fn
        self.sync_timeout_s = sync_timeout_s   # bound of every cross-GPU wait inside the fused decode kernel
        self._symm = None
        self._after_gemm = False   # the last fused call was a T > 16 GEMM (wrote buffer 0 outside the decode parity rule)
        if mm_fn is None:
            from kernels import _ext  # the CUDA library; raises if it is not built (no fallback)
            self._ext = _ext
            self._fmt_id = _ext.FMT_ID[fmt]
        if mode == "fused":
            self._init_symm()
        elif mode != "nccl":
            raise ValueError(mode)

    # ---- fused path: peer-mapped buffers ----
    LL_MAX_TK = 65536   # T * K limit of the fused decode kernel (every CTA of a peer polls all activation lines)

    def _init_symm(self):
        import torch.distributed._symmetric_memory as symm_mem
        dev = self.A.device
        # GEMM path (T > DECODE_MAX_T): the epilogue peer-stores tiles into buffer 0 of every rank; decode path: local only
        self._out = symm_mem.empty((2, self.max_tokens, self.O), dtype=torch.float16, device=dev)   # double-buffered C
        self._symm = symm_mem.rendezvous(self._out, self.group)
        tcap = self.DECODE_MAX_T
        self._xbuf = torch.zeros((2, tcap, self.K), dtype=torch.float16, device=dev)    # rank 0: the step's activations
        # landing buffers of the flag-in-data exchange (16-byte lines carrying 4 fp16): two epoch-parity halves each
        self._x_half = min(tcap * self.K, self.LL_MAX_TK) * 4
        self._xland = symm_mem.empty((2 * self._x_half,), dtype=torch.uint8, device=dev)
        self._xland.zero_()
        self._xlsymm = symm_mem.rendezvous(self._xland, self.group)
        self._c_half = self.world * tcap * self.per * 4
        self._cland = symm_mem.empty((2 * self._c_half,), dtype=torch.uint8, device=dev)
        self._cland.zero_()
        self._clsymm = symm_mem.rendezvous(self._cland, self.group)
        self._counter = torch.zeros(1, dtype=torch.int32, device=dev)
        self._epoch_dev = torch.zeros(1, dtype=torch.int32, device=dev)   # kernel-maintained epoch (replayable mode)
        self._epoch = 0                                                   # host mirror of it
        self._status = torch.zeros(1, dtype=torch.int32, device=dev)      # GGQ_SYNC_* code of a wait that gave up
        torch.cuda.synchronize(dev)
        self._symm.barrier(channel=0)  # everyone's landing buffers are zeroed before anyone sends
        self._prepare_fast_path()

    def _out_ptrs(self, cur: int) -> list[int]:
        off = (cur * self.max_tokens * self.O + self.lo) * 2   # buffer `cur`, this rank's column offset (bytes)
        ptrs = [int(p) + off for p in self._symm.buffer_ptrs]
        return [ptrs[self.rank]] + [p for i, p in enumerate(ptrs) if i != self.rank]

    def _prepare_fast_path(self):
        """Everything about a fused decode step is fixed here, once: a step is one ctypes call with the same arguments
        every time (the kernel keeps the exchange epoch), which is also what makes it CUDA-graph capturable."""
        import ctypes
        ext = self._ext
        self._lib = ext.lib()
        ext.bind_mm_sync(self._lib)
        self._c_local = [self._out[c].data_ptr() for c in (0, 1)]          # column 0 of this rank's [T, O] results
        self._x_local = [self._xbuf[c].data_ptr() for c in (0, 1)]
        sync = ext.PeerSync()
        sync.rank, sync.world, sync.x_owner = self.rank, self.world, 0
        sync.counter = self._counter.data_ptr()
        # replayable mode: the kernel keeps the epoch, even epochs use buffer set 0, odd epochs buffer set 1; nothing
        # in the call changes from step to step, so a step can be captured in a CUDA graph
        sync.epoch_dev = self._epoch_dev.data_ptr()
        sync.X_alt = self._x_local[1]
        sync.C_alt = self._c_local[1]
        sync.x_land = self._xland.data_ptr()
        sync.c_land = self._cland.data_ptr()
        sync.x_land_half, sync.c_land_half = self._x_half, self._c_half
        for r in range(self.world):
            sync.x_land_peer[r] = int(self._xlsymm.buffer_ptrs[r])
            sync.c_land_peer[r] = int(self._clsymm.buffer_ptrs[r])
        sync.status = self._status.data_ptr()
        sync.timeout_ns = int(self.sync_timeout_s * 1e9)
        self._sync = sync
        self._sync_ref = ctypes.byref(sync)
        self._ctas = ctypes.c_int(0)
        self._ctas_ref = ctypes.byref(self._ctas)
        self._a_ptr = self.A.data_ptr()

    def fused_decode_ok(self, T: int) -> bool:
        """Can T tokens take the one-kernel fused decode step?"""
        return (self.mode == "fused" and self.world > 1 and T <= self.DECODE_MAX_T and T * self.K <= self.LL_MAX_TK
                and self.per >= 16 and self.per % 4 == 0 and self.K % 4 == 0)

    def input_buffer(self, T: int) -> torch.Tensor:
        """Rank 0: the [T, K] buffer the NEXT fused decode step reads its activations from.  Writing the activations
        there directly (e.g. as the H2D copy target) and calling forward(None, T=T) skips the staging copy."""
        return self._xbuf[(self._epoch + 1) & 1, :T]

    def set_resident_input(self, X: torch.Tensor) -> None:
        """Rank 0: keep the same activations resident in both slots (benchmarking `forward(None, T=T)`)."""
        self._xbuf[:, :X.shape[0]].copy_(X)

    def _launch_sync(self, T: int) -> None:
        """One fused step with the activations in rank 0's buffer (static arguments: graph-capturable)."""
        rc = self._lib.ggq_mm_sync(self._fmt_id, self._a_ptr, self._x_local[0], self.K, self._c_local[0], self.O,
                                   self.per, T, self.K, self._sync_ref, self._ctas_ref,
                                   torch.cuda.current_stream().cuda_stream)
        if rc != 0:
            self._ext.check(rc, "ggq_mm_sync")

    def _forward_fused_decode(self, X, T: int, broadcast: bool) -> torch.Tensor:
        if not (broadcast and self.world > 1):
            raise ValueError("the fused decode path takes its activations from rank 0 (broadcast=True, world > 1)")
        self._epoch += 1
        cur = self._epoch & 1
        if self.rank == 0 and X is not None and X.data_ptr() != self._x_local[cur]:
            self._xbuf[cur, :T].copy_(X)              # stream-ordered: lands before the kernel below starts
        self._launch_sync(T)
        return self._out[cur, :T]

    def capture_steps(self, T: int, n: int):
        """CUDA graph of `n` consecutive fused decode steps on the activations resident in rank 0's buffer
        (set_resident_input / input_buffer).  Returns replay(): every rank must replay the same number of times;
        last_output(T) is the result of the last step."""
        if not self.fused_decode_ok(T):
            raise ValueError("capture_steps needs the fused decode path")
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for _ in range(n):
                self._launch_sync(T)

        def replay():
            if self._after_gemm:
                # same rule as forward(): a T > 16 GEMM step wrote buffer 0 outside the even/odd discipline
                self._symm.barrier(channel=0)
                self._after_gemm = False
            g.replay()
            self._epoch += n
        return replay

    def sync_status(self) -> int:
        """0, or the GGQ_SYNC_* code of the first cross-GPU wait of a fused decode step that gave up after
        `sync_timeout_s` (1: the activations never arrived, 2: a peer's epoch flag never arrived — a rank did not make
        the matching call).  Synchronises the stream."""
        return int(self._status.item()) if self.mode == "fused" else 0

    DMA_MIN_BYTES = 1 << 20   # slices at least this large are exchanged by the copy engines instead of epilogue peer stores

    def _push_columns(self, src: int, dsts: list[int], T: int) -> None:
        """One 2-D DMA copy per peer, each on its own side stream (several copy engines at once); the current stream
        continues when all of them are done."""
        import ctypes
        L = self._ext.lib()
        if not getattr(L, "_push_bound", False):
            L.ggq_push_columns.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, ctypes.c_int64,
                                           ctypes.c_int64, ctypes.c_int64, ctypes.c_void_p]
            L.ggq_push_columns.restype = ctypes.c_int
            L._push_bound = True
        if getattr(self, "_push_streams", None) is None:
            self._push_streams = [torch.cuda.Stream(device=self.A.device) for _ in dsts]
            self._push_done = [torch.cuda.Event() for _ in dsts]
            self._gemm_done = torch.cuda.Event()
        main = torch.cuda.current_stream()
        self._gemm_done.record(main)
        for st, ev, dst in zip(self._push_streams, self._push_done, dsts):
            st.wait_event(self._gemm_done)
            one = (ctypes.c_void_p * 1)(dst)
            rc = L.ggq_push_columns(src, one, 1, self.O * 2, self.per * 2, T, st.cuda_stream)
            self._ext.check(rc, "ggq_push_columns")
            ev.record(st)
            main.wait_event(ev)

    def _src_rank(self) -> int:
        """Global rank of the group's rank 0 (dist.broadcast takes global ranks)."""
        return 0 if self.group is dist.group.WORLD else dist.get_global_rank(self.group, 0)

    def last_output(self, T: int) -> torch.Tensor:
        """The [T, O] result of the most recent fused step (a T > 16 GEMM step writes buffer 0, a decode step the buffer of
        its epoch's parity)."""
        return self._out[0 if self._after_gemm else self._epoch & 1, :T]

    def forward(self, X, *, broadcast: bool = True, T: int | None = None) -> torch.Tensor:
        """X: fp16 [T, K] (valid on rank 0 when `broadcast`; None + T = the activations were written into
        input_buffer(T), fused decode path only).  Returns C[T, O] on every rank."""
        T = X.shape[0] if X is not None else T
        if self.mode == "fused":
            if T > self.max_tokens:
                raise ValueError(f"T={T} exceeds max_tokens={self.max_tokens} of the symmetric buffer")
            if self.fused_decode_ok(T) and broadcast:
                if self._after_gemm:
                    # the previous call wrote buffer 0 outside the decode path's even/odd discipline: every rank must be
                    # done reading that result before any rank's decode kernel peer-stores into buffer 0 again
                    self._symm.barrier(channel=0)
                    self._after_gemm = False
                return self._forward_fused_decode(X, T, broadcast)
            if X is None:
                raise ValueError("forward(None, T=...) is the fused decode path only (fused_decode_ok(T))")
            if broadcast and self.world > 1:
                dist.broadcast(X, src=self._src_rank(), group=self.group)
            self._symm.barrier(channel=0)   # every rank has finished reading the previous result
            ptrs = self._out_ptrs(0)        # own slice first, then the peers' (same columns of their buffers)
            if T * self.per * 2 >= self.DMA_MIN_BYTES:
                # prefill-sized slice: GEMM into the own buffer, then one 2-D DMA copy per peer (large NVLink transfers by
                # the copy engines; peers in ring order so that no rank is everybody's first destination)
                self._ext.mm_ex(self._fmt_id, self.A, X, ptrs[:1], self.O, self.per, T, self.K)
                peers = ptrs[1:]
                k = self.rank % max(1, len(peers))
                self._push_columns(ptrs[0], peers[k:] + peers[:k], T)
            else:
                self._ext.mm_ex(self._fmt_id, self.A, X, ptrs, self.O, self.per, T, self.K)   # tiles peer-stored by the epilogue
            self._symm.barrier(channel=1)   # every rank's slice has landed everywhere
            self._after_gemm = True
            return self._out[0, :T]
        if broadcast and self.world > 1:
            dist.broadcast(X, src=self._src_rank(), group=self.group)
        if self.mm_fn is not None:
            c = self.mm_fn(self.A, X, self.per, T, self.K)
        else:
            c = self._ext.mm(self._fmt_id, self.A, X, self.per, T, self.K)
        if self.world == 1:
            return c
        gathered = torch.empty((self.world * T, self.per), dtype=c.dtype, device=c.device)
        dist.all_gather_into_tensor(gathered, c.contiguous(), group=self.group)  # rank-major concatenation
        return assemble(gathered.view(self.world, T, self.per))
