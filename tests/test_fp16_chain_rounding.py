"""The vectorised reference-arithmetic kernels (csrc/refmode.cu: acc16h, HADD2 chains) advance the fp16 accumulator with ONE
half-precision add, while torch's CPU kernel — what kernels/cpu_impls/mmq_*_q8_1_cpu.py run (`C += r`, fp16 tensors) — and
the oracle port add in fp32 and round the sum to fp16.  The two agree bit for bit because a second rounding from a format
with P >= 2p + 2 significand bits (fp32: 24, fp16: 11) is innocuous for +, -, *, /.  Checked here against the exact sum: two
fp16 values add exactly in float64 (their significands span at most 11 + 40 bits), and float64 -> fp16 rounds once."""
import numpy as np


def _pairs():
    rng = np.random.default_rng(5)
    bits = rng.integers(0, 1 << 16, size=(2, 4_000_000), dtype=np.uint16)
    a, b = bits[0].view(np.float16), bits[1].view(np.float16)
    # structured cases: near-ties (b a fraction of a's ulp), subnormals, the largest finite values, cancellations
    base = rng.integers(0, 1 << 16, size=200_000, dtype=np.uint16).view(np.float16)
    shift = rng.integers(1, 14, size=base.size)
    with np.errstate(invalid="ignore", over="ignore"):
        tiny = (np.abs(base.astype(np.float64)) * 2.0 ** -(10 + shift) * rng.choice([1.0, 1.5, 0.5, 0.75], size=base.size)).astype(np.float16)
    edge = np.array([0x0001, 0x0002, 0x03FF, 0x0400, 0x7BFF, 0x7BFE, 0x3C00, 0x3C01, 0x8001, 0xFBFF, 0x0000, 0x8000], dtype=np.uint16).view(np.float16)
    ea, eb = np.meshgrid(edge, edge)
    with np.errstate(invalid="ignore", over="ignore"):
        a = np.concatenate([a, base, base, ea.ravel()])
        b = np.concatenate([b, tiny, -base + tiny, eb.ravel()])
    ok = np.isfinite(a) & np.isfinite(b)
    return a[ok], b[ok]


def test_fp32_add_then_round_equals_correctly_rounded_fp16_add():
    a, b = _pairs()
    with np.errstate(over="ignore"):
        twice = (a.astype(np.float32) + b.astype(np.float32)).astype(np.float16)    # torch CPU / oracle / byte-wise kernel
        once = (a.astype(np.float64) + b.astype(np.float64)).astype(np.float16)     # exact sum, one rounding = HADD
    assert np.array_equal(twice.view(np.uint16), once.view(np.uint16))


def test_fp16_product_rounds_once_through_fp32():
    """Q8_0's block scale fp16(d_w * d_x): the fp32 product of two fp16 values is exact (22 significand bits)."""
    a, b = _pairs()
    with np.errstate(over="ignore", under="ignore"):
        p32 = a.astype(np.float32) * b.astype(np.float32)
        p64 = a.astype(np.float64) * b.astype(np.float64)
    assert np.array_equal(p32.astype(np.float64), p64)
