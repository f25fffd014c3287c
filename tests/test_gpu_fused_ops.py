"""GPU tests of the rows next to the mmq path (SURVEY §8f-4 and the boundary): the fused SwiGLU up-projection, the
host-buffer pipe and the two bindings of the C ABI."""
import numpy as np
import pytest
import torch

from oracle import ggq_oracle as orc

pytestmark = pytest.mark.gpu
FMTS = ("q8_0", "q4_k", "q6_k")


@pytest.fixture(scope="module")
def ext():
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    from kernels import _ext
    _ext.lib()
    _ext.torch_ext()
    return _ext


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to("cuda:0")


def rand_x(N, K, seed, scale=1.0):
    return (np.random.default_rng(seed).standard_normal((N, K)) * scale).astype(np.float16)


def swiglu_ref(fmt, Ag, Au, X, M, N, K):
    """fp32-accumulated projections (oracle.ref32), rounded to fp16 as two mmq calls return them, silu * up in fp32."""
    g = orc.ref32(fmt, Ag, X, M, N, K).astype(np.float16).astype(np.float32)
    u = orc.ref32(fmt, Au, X, M, N, K).astype(np.float16).astype(np.float32)
    return g / (1.0 + np.exp(-g)) * u


# O % 8 != 0 (ragged last tile), one tile, cluster split-K (T=16 at K=8192), GEMV (T=1), two n-tiles (T>8)
SWIGLU_CASES = [(64, 256, 1), (1004, 2048, 3), (4096, 4096, 1), (14336, 4096, 8), (2048, 8192, 16), (520, 4096, 9),
                (8, 512, 2), (28672, 8192, 1)]


@pytest.mark.parametrize("fmt", FMTS)
@pytest.mark.parametrize("M,K,N", SWIGLU_CASES)
def test_swiglu_fused_decode(ext, fmt, M, K, N):
    from kernels.swiglu import mmq_swiglu
    Ag = orc.random_blocks(fmt, M, K, seed=3)
    Au = orc.random_blocks(fmt, M, K, seed=4)
    X = rand_x(N, K, 7, 0.05)   # keeps silu(gate) * up inside the fp16 range for random blocks
    L = ext.lib()
    import ctypes
    L.ggq_mm_swiglu_workspace.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64]
    L.ggq_mm_swiglu_workspace.restype = ctypes.c_int64
    fused = L.ggq_mm_swiglu_workspace(ext.FMT_ID[fmt], M, N, K) == 0
    n0 = ext.launch_count()
    C = mmq_swiglu(fmt, dev(Ag), dev(Au), dev(X), M, N, K)
    torch.cuda.synchronize()
    assert ext.launch_count() - n0 == (1 if fused else 3)
    assert C.shape == (N, M) and C.dtype == torch.float16
    got = C.float().cpu().numpy()
    assert np.all(np.isfinite(got))
    rows = np.arange(M) if M * K <= (1 << 24) else np.unique(np.concatenate(
        [np.arange(24), np.arange(M - 24, M), np.random.default_rng(1).choice(M, 200, replace=False)]))
    rb = orc.packed_nbytes(fmt, 1, K)
    sub = lambda A: np.ascontiguousarray(A.reshape(M, rb)[rows]).reshape(-1)  # noqa: E731
    ref = swiglu_ref(fmt, sub(Ag), sub(Au), X, len(rows), N, K)
    mx, fro = orc.tier1_errors(got[:, rows], ref)
    assert mx <= orc.TIER1_MAX and fro <= orc.TIER1_FRO, (fmt, M, K, N, mx, fro)
    # and against the composition of two mmq calls the fusion replaces
    g = ext.mm(ext.FMT_ID[fmt], dev(Ag), dev(X), M, N, K).float()
    u = ext.mm(ext.FMT_ID[fmt], dev(Au), dev(X), M, N, K).float()
    comp = (torch.nn.functional.silu(g) * u).half().float().cpu().numpy()
    mx2, fro2 = orc.tier1_errors(got, comp)
    assert mx2 <= 2e-3 and fro2 <= 1e-3, (fmt, M, K, N, mx2, fro2)


@pytest.mark.parametrize("fmt", FMTS)
def test_swiglu_composed_form_for_many_tokens(ext, fmt):
    """T > 16: gate GEMM -> workspace, up GEMM, one elementwise pass (3 launches), same value."""
    from kernels.swiglu import mmq_swiglu
    M, K, N = 512, 2048, 40
    Ag = orc.random_blocks(fmt, M, K, seed=5)
    Au = orc.random_blocks(fmt, M, K, seed=6)
    X = rand_x(N, K, 8, 0.05)
    n0 = ext.launch_count()
    C = mmq_swiglu(fmt, dev(Ag), dev(Au), dev(X), M, N, K)
    torch.cuda.synchronize()
    assert ext.launch_count() - n0 == 3
    mx, fro = orc.tier1_errors(C.float().cpu().numpy(), swiglu_ref(fmt, Ag, Au, X, M, N, K))
    assert mx <= orc.TIER1_MAX and fro <= orc.TIER1_FRO, (fmt, mx, fro)


@pytest.mark.parametrize("fmt", FMTS)
def test_host_pipe_equals_device_call(ext, fmt):
    """ggq_mm_host: pinned host X in, host C out, rotating slots; every step equals the device-tensor call bit for bit."""
    from kernels.host import HostPipe
    M, K = 1536, 2048
    A = dev(orc.random_blocks(fmt, M, K, seed=9))
    pipe = HostPipe(fmt, A, M, K, max_tokens=8, depth=2)
    xs = [torch.from_numpy(rand_x(1 + (i % 8), K, 20 + i)).pin_memory() for i in range(7)]
    outs = [pipe(x) for x in xs]
    pipe.sync()
    for x, c in zip(xs, outs):
        want = ext.mm(ext.FMT_ID[fmt], A, x.to("cuda:0"), M, x.shape[0], K)
        torch.cuda.synchronize()
        assert torch.equal(want.cpu(), c), fmt
    with pytest.raises(ValueError):
        pipe(torch.zeros((9, K), dtype=torch.float16))
    # pageable host memory works too (the copies are then synchronous)
    x = torch.from_numpy(rand_x(2, K, 99))
    c = pipe(x, out=torch.empty((2, M), dtype=torch.float16))
    pipe.sync()
    assert torch.equal(ext.mm(ext.FMT_ID[fmt], A, x.to("cuda:0"), M, 2, K).cpu(), c)
    pipe.close()


@pytest.mark.parametrize("fmt", FMTS)
def test_both_bindings_reach_the_same_entry_point(ext, fmt):
    M, K, N = 640, 1024, 5
    A = dev(orc.random_blocks(fmt, M, K, seed=2))
    X = dev(rand_x(N, K, 3))
    a = ext.mm(ext.FMT_ID[fmt], A, X, M, N, K)
    b = ext.mm_ctypes(ext.FMT_ID[fmt], A, X, M, N, K)
    out = torch.empty((N, M), device="cuda:0", dtype=torch.float16)
    c = ext.mm(ext.FMT_ID[fmt], A, X, M, N, K, out=out)
    torch.cuda.synchronize()
    assert c.data_ptr() == out.data_ptr()
    assert torch.equal(a, b) and torch.equal(a, c)
