"""Pins the numpy oracle against outputs of the reference itself (tests/golden, made by
tools/make_golden.py from /root/reference).  CPU only.  Everything must match BIT FOR BIT."""
import numpy as np
import pytest

from oracle import ggq_oracle as orc
from oracle import packers

FMTS = ("q8_0", "q4_k", "q6_k")


def _bits(a):
    a = np.ascontiguousarray(a)
    return a.view({2: np.uint16, 4: np.uint32}[a.dtype.itemsize])


@pytest.mark.parametrize("fmt", FMTS)
def test_dequant_matches_reference(golden, fmt):
    for c in golden[fmt]:
        if fmt == "q6_k":  # reference returns fp32 (utils/quantize/q6_k.py:157)
            got = orc.dequantize_q6_k(c["A"], (c["M"], c["K"]))
            assert got.dtype == np.float32
        else:
            got = orc.dequantize(fmt, c["A"], (c["M"], c["K"]))
            assert got.dtype == np.float16
        assert np.array_equal(_bits(got), _bits(c["D"])), (fmt, c["M"], c["K"])


def test_q8_1_packer_matches_reference(golden):
    for fmt in FMTS:
        for c in golden[fmt]:
            assert np.array_equal(orc.quantize_to_q8_1(c["X"]), c["B"])


def test_q8_0_packer_matches_reference(golden):
    for c in golden["q8_0"]:
        assert np.array_equal(orc.quantize_to_q8_0(c["W"]), c["A"])


@pytest.mark.parametrize("fmt", ("q4_k", "q6_k"))
def test_kquant_packer_bytes(golden, fmt):
    if not packers.ref_packers_available():
        pytest.skip("oracle/_ref not built")
    for c in golden[fmt]:
        assert np.array_equal(packers.quantize(fmt, c["W"]), c["A"])
        # threaded row-chunk packing is byte-identical to one call
    W = np.random.default_rng(1).standard_normal((64, 4096)).astype(np.float16)
    assert np.array_equal(packers.quantize(fmt, W, threads=1), packers.quantize(fmt, W, threads=4))


@pytest.mark.parametrize("fmt", FMTS)
def test_mmq_cpu_matches_reference(golden, fmt):
    for c in golden[fmt]:
        got = orc.MMQ_CPU[fmt](c["A"], c["B"], c["M"], c["N"], c["K"])
        assert got.shape == (c["N"], c["M"]) and got.dtype == np.float16
        assert np.array_equal(_bits(got), _bits(c["C"])), (fmt, c["M"], c["N"], c["K"])


@pytest.mark.parametrize("fmt", FMTS)
def test_tiers_are_consistent(golden, fmt):
    """cpu_impls vs the fp32-accumulated reference: the Q8_1 activation noise keeps the
    two oracles ~5e-3 apart; both tiers must agree to that level on many-output cases."""
    for c in golden[fmt]:
        if c["M"] * c["N"] < 16:
            continue
        r32 = orc.ref32(fmt, c["A"], c["X"], c["M"], c["N"], c["K"])
        mx, fro = orc.tier1_errors(c["C"].astype(np.float32), r32)
        assert fro < 2.5e-2 and mx < 5e-2, (fmt, mx, fro)


@pytest.mark.parametrize("fmt", FMTS)
def test_random_blocks_are_valid(fmt):
    A = orc.random_blocks(fmt, 8, 1024, seed=3)
    assert A.dtype == np.int8 and A.size == orc.packed_nbytes(fmt, 8, 1024)
    W = orc.dequantize(fmt, A, (8, 1024))
    assert np.all(np.isfinite(W.astype(np.float32)))
    assert float(np.abs(W.astype(np.float32)).max()) > 0


def test_allclose_ref_semantics():
    b = np.array([1.0, -2.0, 100.0])
    assert orc.allclose_ref(b + 0.9, b, 0.01)
    assert not orc.allclose_ref(b + 1.1, b, 0.01)
    assert not orc.allclose_ref(b, np.array([1.0, np.nan, 2.0]), 0.01)
