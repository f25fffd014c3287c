"""CPU oracle for the GGUF mmq hot path (Q8_0 / Q4_K / Q6_K weights x fp16 activations).

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is part of the product path: it may be
imported only by ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs, and only as the checker / the thing timed as "the CPU path" — the
shipped kernels (``gguf-triton-kernel_b200/``) never import it and have no CPU fallback.

Parity status: PINNED.  Every function here is a numpy restatement of a reference function and is
checked bit-for-bit against outputs of the reference itself (``tests/golden/*.npz``, produced by
``tools/make_golden.py`` which imports the reference from ``/root/reference``); see
``tests/test_oracle_golden.py``.

Reference anchors (all paths relative to the reference repo root):

* block layouts ............ kernels/mmq_q4_k.py:1-16, kernels/mmq_q6_k.py:1-14,
                             utils/quantize/q8_1.py:1-13, utils/quantize/q6_k.py:15-19
* dequantizers ............. utils/quantize/q8_0.py:52-100, q4_k.py:100-158, q6_k.py:117-159
* fp16 packers ............. utils/quantize/q8_0.py:4-49, q8_1.py:18-70
* mmq CPU implementations .. kernels/cpu_impls/mmq_q8_0_q8_1_cpu.py:5-56,
                             mmq_q4_k_q8_1_cpu.py:5-119, mmq_q6_k_q8_1_cpu.py:5-152
* test criterion ........... utils/test_utils.py:4-8

Naming follows the reference: ``A`` = packed weight with ``M`` rows (out-features), ``B`` =
activation with ``N`` rows (tokens), result is ``[N, M]``.
"""
from __future__ import annotations

import numpy as np

QK8_0 = 32
QK8_1 = 32
QK_K = 256
Q8_0_SIZE = 34
Q8_1_SIZE = 36
Q4_K_SIZE = 144
Q6_K_SIZE = 210

#: fmt -> (elements per block, bytes per block)
FORMATS = {"q8_0": (QK8_0, Q8_0_SIZE), "q4_k": (QK_K, Q4_K_SIZE), "q6_k": (QK_K, Q6_K_SIZE)}


def packed_nbytes(fmt: str, rows: int, K: int) -> int:
    qk, blk = FORMATS[fmt]
    assert K % qk == 0
    return rows * (K // qk) * blk


def _as_u8(packed) -> np.ndarray:
    a = np.ascontiguousarray(np.asarray(packed))
    assert a.dtype in (np.int8, np.uint8), a.dtype
    return a.reshape(-1).view(np.uint8)


def _f16(raw_u8_pairs: np.ndarray) -> np.ndarray:
    """(..., 2) uint8 little-endian -> (...) float16 (bitcast)."""
    return np.ascontiguousarray(raw_u8_pairs).view("<f2")[..., 0]


# --------------------------------------------------------------------------------------------
# Unpacking (integer fields) — shared by the dequantizers and the mmq restatements
# --------------------------------------------------------------------------------------------

def unpack_q8_0(packed):
    """-> d[nb] float16, q[nb, 32] int8.   Layout: utils/quantize/q8_0.py:41-47."""
    b = _as_u8(packed).reshape(-1, Q8_0_SIZE)
    return _f16(b[:, 0:2]), b[:, 2:].view(np.int8)


def unpack_q8_1(packed):
    """-> d[nb] float16, s[nb] float16, q[nb, 32] int8.   Layout: utils/quantize/q8_1.py:61-68."""
    b = _as_u8(packed).reshape(-1, Q8_1_SIZE)
    return _f16(b[:, 0:2]), _f16(b[:, 2:4]), b[:, 4:].view(np.int8)


def unpack_q4_k(packed):
    """-> d[nb] f16, dmin[nb] f16, sc[nb, 8] u8, m[nb, 8] u8, q[nb, 256] u8 (element order).

    6-bit scale/min packing: utils/quantize/q4_k.py:100-122 (same as q4_k_ref.c:174-186 and
    kernels/cpu_impls/mmq_q4_k_q8_1_cpu.py:31-52); nibble order: q4_k.py:139-140,
    mmq_q4_k_q8_1_cpu.py:55-56.
    """
    b = _as_u8(packed).reshape(-1, Q4_K_SIZE)
    d, dmin = _f16(b[:, 0:2]), _f16(b[:, 2:4])
    s = b[:, 4:16]
    lo, mi, hi = s[:, 0:4], s[:, 4:8], s[:, 8:12]
    sc = np.concatenate([lo & 0x3F, (hi & 0x0F) | ((lo >> 2) & 0x30)], axis=1)
    mn = np.concatenate([mi & 0x3F, (hi >> 4) | ((mi >> 2) & 0x30)], axis=1)
    qs = b[:, 16:].reshape(-1, 4, 32)
    q = np.stack([qs & 0x0F, qs >> 4], axis=2).reshape(-1, QK_K)
    return d, dmin, sc, mn, q


def unpack_q6_k(packed):
    """-> d[nb] f16, sc[nb, 16] int8, q[nb, 256] int8 (element order, 32 already subtracted).

    ql/qh interleave: utils/quantize/q6_k.py:128-133, table in kernels/mmq_q6_k.py:39-48 and
    kernels/cpu_impls/mmq_q6_k_q8_1_cpu.py:38-77.
    """
    b = _as_u8(packed).reshape(-1, Q6_K_SIZE)
    ql = b[:, 0:128].reshape(-1, 2, 1, 64)
    qh = b[:, 128:192].reshape(-1, 2, 1, 32)
    sc = b[:, 192:208].view(np.int8)
    d = _f16(b[:, 208:210])
    lo = (ql >> np.array([0, 4], dtype=np.uint8).reshape(1, 1, 2, 1)) & 0x0F  # [nb,2,2,64]
    lo = lo.reshape(-1, 2, 4, 32)  # per half: (ql[0:32]&F, ql[32:64]&F, ql[0:32]>>4, ql[32:64]>>4)
    hi = (qh >> np.array([0, 2, 4, 6], dtype=np.uint8).reshape(1, 1, 4, 1)) & 0x03  # [nb,2,4,32]
    q = (lo | (hi << 4)).astype(np.int8) - np.int8(32)
    return d, sc, q.reshape(-1, QK_K)


# --------------------------------------------------------------------------------------------
# Dequantizers (Tier-0 targets: the CUDA dequant must equal these bit for bit)
# --------------------------------------------------------------------------------------------

def dequantize_q8_0(packed, shape) -> np.ndarray:
    """fp16 * fp16 -> fp16, one rounding.   utils/quantize/q8_0.py:94."""
    d, q = unpack_q8_0(packed)
    return (q.astype(np.float16) * d[:, None]).reshape(shape)


def dequantize_q4_k_f32(packed, shape) -> np.ndarray:
    """fp32 (d*sc)*q - (dmin*m); every product is exact, one rounding at the subtract.
    utils/quantize/q4_k.py:125-143."""
    d, dmin, sc, mn, q = unpack_q4_k(packed)
    ds = d.astype(np.float32)[:, None] * sc.astype(np.float32)
    dm = dmin.astype(np.float32)[:, None] * mn.astype(np.float32)
    qf = q.reshape(-1, 8, 32).astype(np.float32)
    return (ds[:, :, None] * qf - dm[:, :, None]).reshape(shape)


def dequantize_q4_k(packed, shape) -> np.ndarray:
    """utils/quantize/q4_k.py:156 (``.to(torch.float16)`` of the fp32 result)."""
    return dequantize_q4_k_f32(packed, shape).astype(np.float16)


def dequantize_q6_k(packed, shape) -> np.ndarray:
    """Returns fp32 like the reference (utils/quantize/q6_k.py:117-159); exact, no rounding."""
    d, sc, q = unpack_q6_k(packed)
    ds = d.astype(np.float32)[:, None] * sc.astype(np.float32)
    return (ds[:, :, None] * q.reshape(-1, 16, 16).astype(np.float32)).reshape(shape)


def dequantize(fmt: str, packed, shape) -> np.ndarray:
    """Dequantized weights as float16 — the bit-exact target of ``ggq_dequant_*_f16``."""
    if fmt == "q8_0":
        return dequantize_q8_0(packed, shape)
    if fmt == "q4_k":
        return dequantize_q4_k(packed, shape)
    if fmt == "q6_k":
        return dequantize_q6_k(packed, shape).astype(np.float16)
    raise KeyError(fmt)


# --------------------------------------------------------------------------------------------
# fp16 packers implemented in Python by the reference (Q8_0 weights, Q8_1 activations)
# --------------------------------------------------------------------------------------------

def _group_scale_and_q(x):
    x = np.asarray(x, dtype=np.float16).reshape(-1, 32)
    amax = np.max(np.abs(x), axis=1)
    return x, amax


def quantize_to_q8_0(x) -> np.ndarray:
    """All-fp16 arithmetic: d = max|x| / 127 (fp16), q = rint(x / d) clipped to +-127; d = 1 for an
    all-zero group.   utils/quantize/q8_0.py:4-49."""
    g, amax = _group_scale_and_q(x)
    d = np.ones(g.shape[0], dtype=np.float16)
    nz = amax != 0
    d[nz] = amax[nz] / np.float16(127.0)
    with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
        q = np.clip(np.rint(g / d[:, None]), -127, 127).astype(np.int8)
    out = np.empty((g.shape[0], Q8_0_SIZE), dtype=np.int8)
    out[:, 0:2] = d.view(np.int8).reshape(-1, 2)
    out[:, 2:] = q
    return out.reshape(-1)


def quantize_to_q8_1(x) -> np.ndarray:
    """d = max|x| / 127 (0 for an all-zero group), q = rint(x / d), s = d * fp16(sum q).
    utils/quantize/q8_1.py:18-70."""
    g, amax = _group_scale_and_q(x)
    d = np.zeros(g.shape[0], dtype=np.float16)
    nz = amax != 0
    d[nz] = amax[nz] / np.float16(127.0)
    d_safe = d.copy()
    d_safe[d_safe == 0] = np.float16(1.0)
    with np.errstate(over="ignore", invalid="ignore", divide="ignore"):
        q = np.clip(np.rint(g / d_safe[:, None]), -127, 127).astype(np.int8)
    s = d * q.sum(axis=1, dtype=np.int32).astype(np.float16)
    out = np.empty((g.shape[0], Q8_1_SIZE), dtype=np.int8)
    out[:, 0:2] = d.view(np.int8).reshape(-1, 2)
    out[:, 2:4] = s.view(np.int8).reshape(-1, 2)
    out[:, 4:] = q
    return out.reshape(-1)


# --------------------------------------------------------------------------------------------
# mmq CPU implementations (Tier-2 oracle): Q8_1 activations, integer block dots, fp16 accumulator
# --------------------------------------------------------------------------------------------

def _acc16(C16: np.ndarray, r32: np.ndarray) -> np.ndarray:
    """``C[m, n] += r.item()`` on a float16 tensor (mmq_q4_k_q8_1_cpu.py:117): torch first rounds the
    Python scalar to the tensor dtype (fp16), then adds in fp32 and rounds the sum to fp16
    [checked against torch 2.11 on 2000 random pairs while writing this oracle]."""
    r16 = r32.astype(np.float16)
    return (C16.astype(np.float32) + r16.astype(np.float32)).astype(np.float16)


def mmq_q8_0_q8_1_cpu(A, B, M: int, N: int, K: int) -> np.ndarray:
    """kernels/cpu_impls/mmq_q8_0_q8_1_cpu.py:5-56.  Per block pair:
    ``fp16(d_w * d_x)`` (fp16 product), times the int32 dot in fp32, rounded to fp16, then added
    into the fp16 accumulator.  Vectorised over (m, n); block order is sequential as in the
    reference so the fp16 accumulation rounds identically."""
    assert K % 32 == 0
    nb = K // 32
    dA, qA = unpack_q8_0(A)
    dB, _, qB = unpack_q8_1(B)
    assert dA.size == M * nb and dB.size == N * nb
    dA, qA = dA.reshape(M, nb), qA.reshape(M, nb, 32).astype(np.int32)
    dB, qB = dB.reshape(N, nb), qB.reshape(N, nb, 32).astype(np.int32)
    C = np.zeros((M, N), dtype=np.float16)
    with np.errstate(over="ignore", invalid="ignore"):
        for b in range(nb):
            dot = qA[:, b, :] @ qB[:, b, :].T
            sc = dA[:, b, None] * dB[None, :, b]  # fp16 * fp16 -> fp16
            r = (sc.astype(np.float32) * dot.astype(np.float32)).astype(np.float16)
            C = _acc16(C, r.astype(np.float32))
    return C.T


def mmq_q4_k_q8_1_cpu(A, B, M: int, N: int, K: int) -> np.ndarray:
    """kernels/cpu_impls/mmq_q4_k_q8_1_cpu.py:61-119.  Per (super-block, sub-block):
    ``((d*sc)*d_x)*dot - (dmin*m)*s_x`` left to right in fp32, added into the fp16 accumulator."""
    assert K % 256 == 0
    nb = K // 256
    d, dmin, sc, mn, q = unpack_q4_k(A)
    dB, sB, qB = unpack_q8_1(B)
    assert d.size == M * nb and dB.size == N * nb * 8
    d = d.astype(np.float32).reshape(M, nb)
    dmin = dmin.astype(np.float32).reshape(M, nb)
    sc = sc.astype(np.float32).reshape(M, nb, 8)
    mn = mn.astype(np.float32).reshape(M, nb, 8)
    q = q.reshape(M, nb, 8, 32).astype(np.int32)
    dB = dB.astype(np.float32).reshape(N, nb, 8)
    sB = sB.astype(np.float32).reshape(N, nb, 8)
    qB = qB.reshape(N, nb, 8, 32).astype(np.int32)
    C = np.zeros((M, N), dtype=np.float16)
    with np.errstate(over="ignore", invalid="ignore"):
        for b in range(nb):
            for j in range(8):
                dot = (q[:, b, j, :] @ qB[:, b, j, :].T).astype(np.float32)
                t1 = ((d[:, b] * sc[:, b, j])[:, None] * dB[None, :, b, j]) * dot
                t2 = (dmin[:, b] * mn[:, b, j])[:, None] * sB[None, :, b, j]
                C = _acc16(C, t1 - t2)
    return C.T


def mmq_q6_k_q8_1_cpu(A, B, M: int, N: int, K: int) -> np.ndarray:
    """kernels/cpu_impls/mmq_q6_k_q8_1_cpu.py:84-152.  Per Q8_1 block (two 16-wide sub-blocks):
    ``d_x * ((d*sc1)*dot1 + (d*sc2)*dot2)`` in fp32, added into the fp16 accumulator."""
    assert K % 256 == 0
    nb = K // 256
    d, sc, q = unpack_q6_k(A)
    dB, _, qB = unpack_q8_1(B)
    assert d.size == M * nb and dB.size == N * nb * 8
    d = d.astype(np.float32).reshape(M, nb)
    sc = sc.astype(np.float32).reshape(M, nb, 16)
    q = q.reshape(M, nb, 16, 16).astype(np.int32)
    dB = dB.astype(np.float32).reshape(N, nb, 8)
    qB = qB.reshape(N, nb, 8, 2, 16).astype(np.int32)
    C = np.zeros((M, N), dtype=np.float16)
    with np.errstate(over="ignore", invalid="ignore"):
        for b in range(nb):
            for j in range(8):
                dot1 = (q[:, b, 2 * j, :] @ qB[:, b, j, 0, :].T).astype(np.float32)
                dot2 = (q[:, b, 2 * j + 1, :] @ qB[:, b, j, 1, :].T).astype(np.float32)
                s1 = (d[:, b] * sc[:, b, 2 * j])[:, None]
                s2 = (d[:, b] * sc[:, b, 2 * j + 1])[:, None]
                r = dB[None, :, b, j] * (s1 * dot1 + s2 * dot2)
                C = _acc16(C, r)
    return C.T


MMQ_CPU = {"q8_0": mmq_q8_0_q8_1_cpu, "q4_k": mmq_q4_k_q8_1_cpu, "q6_k": mmq_q6_k_q8_1_cpu}


def mmq_cpu(fmt: str, A, X16, M: int, N: int, K: int) -> np.ndarray:
    """The reference's CPU path for fp16 activations: pack them to Q8_1 (as the reference tests do,
    test/test_mmq_q4_k.py:31-33) and run the matching cpu_impl."""
    return MMQ_CPU[fmt](A, quantize_to_q8_1(X16), M, N, K)


# --------------------------------------------------------------------------------------------
# Tier-1 reference and the acceptance criteria
# --------------------------------------------------------------------------------------------

def ref32(fmt: str, A, X16, M: int, N: int, K: int) -> np.ndarray:
    """``X.float() @ dequant(W).float().T`` accumulated in float64 — the fp32-accumulated reference
    the north-star tolerance is stated against.  Returns float32 ``[N, M]``."""
    W = dequantize(fmt, A, (M, K)).astype(np.float64)
    X = np.asarray(X16, dtype=np.float16).reshape(N, K).astype(np.float64)
    return (X @ W.T).astype(np.float32)


def allclose_ref(a, b, atol_ratio: float = 0.01) -> bool:
    """utils/test_utils.py:4-8: ``|a-b| <= atol_ratio*max|b| + 1e-5*|b|``; False when b has NaN."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    mx = np.max(np.abs(b)) if b.size else 0.0
    if np.isnan(mx):
        return False
    return bool(np.all(np.abs(a - b) <= atol_ratio * mx + 1e-5 * np.abs(b)))


def tier1_errors(c, ref) -> tuple[float, float]:
    """(max|err| / max|ref|, ||err||_F / ||ref||_F)."""
    c = np.asarray(c, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    err = c - ref
    mx = float(np.max(np.abs(ref))) if ref.size else 0.0
    fro = float(np.linalg.norm(ref))
    return (float(np.max(np.abs(err))) / mx if mx else 0.0,
            float(np.linalg.norm(err)) / fro if fro else 0.0)


TIER1_MAX = 1e-2   # max|err| <= 1e-2 * max|ref|
TIER1_FRO = 2e-3   # ||err||_F <= 2e-3 * ||ref||_F


def tier1_ok(c, ref) -> bool:
    mx, fro = tier1_errors(c, ref)
    return mx <= TIER1_MAX and fro <= TIER1_FRO


# --------------------------------------------------------------------------------------------
# Synthetic packed weights: every byte pattern with finite fp16 scales is a valid block
# --------------------------------------------------------------------------------------------

def random_blocks(fmt: str, rows: int, K: int, seed: int = 0, scale: float = 0.02) -> np.ndarray:
    """Random *valid* packed weights, flat int8 of ``packed_nbytes(fmt, rows, K)`` bytes.  Quant
    payload bytes are uniform random (all code points incl. -128 / 63 / 0xFF get exercised);
    fp16 super-scales are finite, ``|d| ~ U(0.25, 1)*scale`` with random sign for d (Q8_0, Q6_K
    allow it) and non-negative d/dmin for Q4_K as its packer produces."""
    qk, blk = FORMATS[fmt]
    nb = rows * (K // qk)
    rng = np.random.default_rng(seed)
    raw = rng.integers(0, 256, size=(nb, blk), dtype=np.uint8)

    def scales(n, signed):
        v = rng.uniform(0.25, 1.0, size=n) * scale
        if signed:
            v *= rng.choice([-1.0, 1.0], size=n)
        return v.astype(np.float16).view(np.uint8).reshape(n, 2)

    if fmt == "q8_0":
        raw[:, 0:2] = scales(nb, True)
    elif fmt == "q4_k":
        raw[:, 0:2] = scales(nb, False) if scale else 0
        raw[:, 2:4] = scales(nb, False)
        # keep d*sc*q (<= 63*15*d) comfortably inside fp16: scale/16
        raw[:, 0:2] = (raw[:, 0:2].copy().view("<f2").astype(np.float32) / 16).astype(np.float16).view(np.uint8).reshape(nb, 2)
        raw[:, 2:4] = (raw[:, 2:4].copy().view("<f2").astype(np.float32) / 16).astype(np.float16).view(np.uint8).reshape(nb, 2)
    elif fmt == "q6_k":
        raw[:, 208:210] = scales(nb, True)
        raw[:, 208:210] = (raw[:, 208:210].copy().view("<f2").astype(np.float32) / 64).astype(np.float16).view(np.uint8).reshape(nb, 2)
    else:
        raise KeyError(fmt)
    return raw.reshape(-1).view(np.int8)
