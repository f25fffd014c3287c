"""GPU mirrors of the reference's utils/quantize functions next to the mmq path: packers must be
byte-identical, dequantizers bit-identical, to the oracle (which is pinned to the reference)."""
import numpy as np
import pytest
import torch

from oracle import ggq_oracle as orc

pytestmark = pytest.mark.gpu


def _x(shape, seed, scale=1.0):
    x = (np.random.default_rng(seed).standard_normal(shape) * scale).astype(np.float16)
    x.reshape(-1)[:64] = 0          # all-zero groups (d = 1 for Q8_0, d = 0 for Q8_1)
    x.reshape(-1)[64:67] = [np.float16(0.5), np.float16(-0.5), np.float16(1.5)]  # ties for round-half-even
    return x


@pytest.mark.parametrize("shape,scale", [((4, 64), 1.0), ((33, 1024), 0.01), ((7, 4096), 30.0)])
def test_q8_packers_byte_identical(shape, scale):
    from utils.quantize.q8_0 import quantize_to_q8_0
    from utils.quantize.q8_1 import quantize_to_q8_1
    x = _x(shape, 5, scale)
    xd = torch.from_numpy(x).cuda()
    assert np.array_equal(quantize_to_q8_0(xd).cpu().numpy(), orc.quantize_to_q8_0(x))
    assert np.array_equal(quantize_to_q8_1(xd).cpu().numpy(), orc.quantize_to_q8_1(x))
    with pytest.raises(ValueError):
        quantize_to_q8_0(torch.zeros(33, dtype=torch.float16, device="cuda"))


def test_golden_q8_packers(golden):
    from utils.quantize.q8_0 import quantize_to_q8_0
    from utils.quantize.q8_1 import quantize_to_q8_1
    for c in golden["q8_0"]:
        assert np.array_equal(quantize_to_q8_0(torch.from_numpy(c["W"]).cuda()).cpu().numpy(), c["A"])
        assert np.array_equal(quantize_to_q8_1(torch.from_numpy(c["X"]).cuda()).cpu().numpy(), c["B"])


def test_dequantize_mirrors(golden):
    from utils.quantize.q4_k import dequantize_q4_k
    from utils.quantize.q6_k import dequantize_q6_k
    from utils.quantize.q8_0 import dequantize_q8_0
    for fmt, fn in (("q8_0", dequantize_q8_0), ("q4_k", dequantize_q4_k), ("q6_k", dequantize_q6_k)):
        for c in golden[fmt]:
            got = fn(torch.from_numpy(c["A"]).cuda(), (c["M"], c["K"])).cpu().numpy()
            assert got.dtype == c["D"].dtype and got.shape == c["D"].shape
            bits = np.uint32 if got.dtype == np.float32 else np.uint16
            assert np.array_equal(got.view(bits), c["D"].view(bits)), fmt


def test_reference_style_pipeline_on_gpu():
    """quantize (GPU) -> mmq (GPU) as in test/test_mmq_q8_0.py:27-36, everything resident on the device."""
    from kernels.mmq_q8_0 import mmq_q8_0
    from utils.quantize.q8_0 import quantize_to_q8_0
    M, N, K = 64, 4, 1024
    W = torch.randn(M, K, dtype=torch.float16, device="cuda")
    X = torch.randn(N, K, dtype=torch.float16, device="cuda")
    A = quantize_to_q8_0(W)
    C = mmq_q8_0(A, X, M, N, K)
    ref = orc.ref32("q8_0", A.cpu().numpy(), X.cpu().numpy(), M, N, K)
    mx, fro = orc.tier1_errors(C.float().cpu().numpy(), ref)
    assert mx <= orc.TIER1_MAX and fro <= orc.TIER1_FRO


@pytest.mark.parametrize("fmt", ("q8_0", "q4_k", "q6_k"))
def test_reference_arithmetic_mode_is_bit_identical(golden, fmt):
    """GPU Q8_1 mode == the reference's kernels/cpu_impls outputs (golden) and == the oracle on fresh inputs."""
    from kernels import q8_1_mode
    from utils.quantize.q8_1 import quantize_to_q8_1
    fn = {"q8_0": q8_1_mode.mmq_q8_0_q8_1, "q4_k": q8_1_mode.mmq_q4_k_q8_1, "q6_k": q8_1_mode.mmq_q6_k_q8_1}[fmt]
    for c in golden[fmt]:
        got = fn(torch.from_numpy(c["A"]).cuda(), torch.from_numpy(c["B"]).cuda(), c["M"], c["N"], c["K"]).cpu().numpy()
        assert np.array_equal(got.view(np.uint16), np.ascontiguousarray(c["C"]).view(np.uint16)), (fmt, c["M"], c["N"], c["K"])
    M, N, K = 48, 5, 2048
    A = orc.random_blocks(fmt, M, K, seed=8)
    X = np.random.default_rng(9).standard_normal((N, K)).astype(np.float16)
    Bq = quantize_to_q8_1(torch.from_numpy(X).cuda())               # GPU packer feeding the GPU parity kernel
    got = fn(torch.from_numpy(A).cuda(), Bq, M, N, K).cpu().numpy()
    want = np.ascontiguousarray(orc.mmq_cpu(fmt, A, X, M, N, K))
    assert np.array_equal(got.view(np.uint16), want.view(np.uint16))


@pytest.mark.parametrize("fmt", ("q8_0", "q4_k", "q6_k"))
@pytest.mark.parametrize("M,N,K", [(300, 3, 4096), (33, 1, 512), (70, 2, 768), (16, 17, 1280),
                                   (257, 8, 1024), (130, 9, 256), (128, 16, 2048)])
def test_reference_arithmetic_mode_both_kernels(fmt, M, N, K):
    """The vectorised kernel (rows that are whole words: K/QK even, always for Q4_K) and the byte-wise one (K = 768, 1280:
    odd block counts for Q6_K / Q4_K-sized super-blocks) are both bit-identical to the oracle's kernels/cpu_impls port;
    M not a multiple of 32 mixes tokens inside a warp.  Q4_K with N >= 2 runs the tensor-core (IMMA) kernel: 8-token tiles
    (N = 9, 17: a last tile of one token; N = 3, 5: a partial tile), row tiles with a ragged end, a single-stage K = 256."""
    from kernels import q8_1_mode
    from utils.quantize.q8_1 import quantize_to_q8_1
    fn = {"q8_0": q8_1_mode.mmq_q8_0_q8_1, "q4_k": q8_1_mode.mmq_q4_k_q8_1, "q6_k": q8_1_mode.mmq_q6_k_q8_1}[fmt]
    A = orc.random_blocks(fmt, M, K, seed=M + N)
    X = np.random.default_rng(K).standard_normal((N, K)).astype(np.float16)
    Bq = quantize_to_q8_1(torch.from_numpy(X).cuda())
    got = fn(torch.from_numpy(A).cuda(), Bq, M, N, K).cpu().numpy()
    want = np.ascontiguousarray(orc.mmq_cpu(fmt, A, X, M, N, K))
    assert np.array_equal(got.view(np.uint16), want.view(np.uint16)), (fmt, M, N, K)


@pytest.mark.parametrize("M,N,K", [(200, 2, 1024), (140, 4, 512), (129, 8, 2048), (64, 11, 768)])
def test_reference_arithmetic_mode_q4k_fallback_kernels(M, N, K):
    """Q4_K kernel selection depends on pointer alignment: Q8_1 activations that are only 4-byte aligned take the
    thread-per-row tile kernel (token tiles of 4 / 8) instead of the tensor-core one, weights that are not 16-byte aligned
    the byte-wise kernel.  All of them are bit-identical to the oracle's port of kernels/cpu_impls."""
    from kernels import q8_1_mode
    from utils.quantize.q8_1 import quantize_to_q8_1
    A = orc.random_blocks("q4_k", M, K, seed=3 * M + N)
    X = np.random.default_rng(K + N).standard_normal((N, K)).astype(np.float16)
    want = np.ascontiguousarray(orc.mmq_cpu("q4_k", A, X, M, N, K)).view(np.uint16)
    Ag = torch.from_numpy(A).cuda()
    Bq = quantize_to_q8_1(torch.from_numpy(X).cuda()).reshape(-1)
    for a_off, b_off in ((0, 0), (0, 4), (0, 8), (2, 0)):
        Abuf = torch.empty(Ag.numel() + 16, dtype=torch.int8, device="cuda")
        Bbuf = torch.empty(Bq.numel() + 16, dtype=torch.int8, device="cuda")
        Av, Bv = Abuf[a_off:a_off + Ag.numel()], Bbuf[b_off:b_off + Bq.numel()]
        Av.copy_(Ag.reshape(-1))
        Bv.copy_(Bq)
        assert Av.data_ptr() % 16 == a_off and Bv.data_ptr() % 16 == b_off
        got = q8_1_mode.mmq_q4_k_q8_1(Av, Bv, M, N, K).cpu().numpy()
        assert np.array_equal(got.view(np.uint16), want), (M, N, K, a_off, b_off)


def _kq_cases():
    rng = np.random.default_rng(1)
    n = 256 * 512
    return {
        "randn": rng.standard_normal(n).astype(np.float32),
        "uniform positive": rng.uniform(0, 3, n).astype(np.float32),
        "tiny": (rng.standard_normal(n) * 1e-20).astype(np.float32),
        "large": (rng.standard_normal(n) * 3e4).astype(np.float32),
        "zeros and constants": np.concatenate([np.zeros(512), np.full(512, 0.37), np.full(512, -1.5)]).astype(np.float32),
        "sparse": np.where(rng.random(n) < 0.1, rng.standard_normal(n), 0).astype(np.float32),
        "mixed scales": (rng.standard_normal(n) * np.repeat(10.0 ** rng.uniform(-6, 3, n // 32), 32)).astype(np.float32),
    }


@pytest.mark.parametrize("name", list(_kq_cases()))
def test_kquant_packers_byte_identical_to_reference_library(name):
    """GPU Q4_K / Q6_K packers vs the reference's compiled packers (oracle/_ref travels with the repo) on raw fp32."""
    import ctypes
    import os
    from utils.quantize.q4_k import quantize_to_q4_k
    from utils.quantize.q6_k import quantize_to_q6_k
    ref_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref")
    x = _kq_cases()[name]
    xd = torch.from_numpy(x).cuda()
    for fn, lib, sym, blk in ((quantize_to_q4_k, "libq4_k_ref.so", "quantize_row_q4_K_ref", 144),
                              (quantize_to_q6_k, "libq6_k_ref.so", "quantize_row_q6_K_ref", 210)):
        ref = np.zeros(x.size // 256 * blk, dtype=np.uint8)
        getattr(ctypes.CDLL(os.path.join(ref_dir, lib)), sym)(x.ctypes.data_as(ctypes.c_void_p), ref.ctypes.data_as(ctypes.c_void_p),
                                                              ctypes.c_longlong(x.size))
        got = fn(xd).cpu().numpy().view(np.uint8)
        assert got.shape == ref.shape and np.array_equal(got, ref), (name, sym, int((got != ref).sum()))


def test_kquant_packers_reproduce_the_golden_fixtures(golden):
    """fp16 weights of the golden cases -> exactly the packed bytes the reference produced for them."""
    from utils.quantize.q4_k import quantize_to_q4_k
    from utils.quantize.q6_k import quantize_to_q6_k
    for fmt, fn in (("q4_k", quantize_to_q4_k), ("q6_k", quantize_to_q6_k)):
        for c in golden[fmt]:
            assert np.array_equal(fn(torch.from_numpy(c["W"]).cuda()).cpu().numpy(), c["A"]), (fmt, c["M"], c["K"])
    with pytest.raises(ValueError):
        quantize_to_q4_k(torch.zeros(100, dtype=torch.float16, device="cuda"))


def test_packed_on_gpu_then_multiplied():
    """pack on the GPU -> mmq on the GPU: the whole synthetic pipeline of the reference's tests without the host."""
    from kernels.mmq_q4_k import mmq_q4_k
    from utils.quantize.q4_k import dequantize_q4_k, quantize_to_q4_k
    O, K, T = 512, 2048, 3
    W = torch.randn((O, K), device="cuda", dtype=torch.float16)
    X = torch.randn((T, K), device="cuda", dtype=torch.float16)
    A = quantize_to_q4_k(W)
    C = mmq_q4_k(A, X, O, T, K)
    ref = X.float() @ dequantize_q4_k(A, (O, K)).float().t()
    err = (C.float() - ref).norm() / ref.norm()
    assert err < 2e-3, float(err)
