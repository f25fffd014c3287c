"""CPU-only checks of the drop-in boundary: libggq.so loads, exports every symbol include/ggq.h
declares, answers its pure-host queries, and the Python entry points keep the reference's names,
constants and error behaviour.  No compute calls (no GPU here)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gguf-triton-kernel_b200", "libggq.so")


def _declared():
    src = open(os.path.join(ROOT, "include", "ggq.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(ggq_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(LIB), "libggq.so not built (make -C gguf-triton-kernel_b200)"
    lib = ctypes.CDLL(LIB)
    names = _declared()
    assert len(names) >= 12
    for n in names:
        assert hasattr(lib, n), f"{n} declared in ggq.h but not exported"


def test_host_queries():
    from kernels import _ext
    L = _ext.lib()
    assert L.ggq_version() == 108
    assert L.ggq_packed_nbytes(0, 4096, 4096) == 17825792          # BASELINE config 1
    assert L.ggq_packed_nbytes(1, 128256, 4096) == 295501824       # config 2
    assert L.ggq_packed_nbytes(2, 4096, 14336) == 48168960         # config 3
    assert L.ggq_packed_nbytes(1, 4, 100) == -1 and L.ggq_packed_nbytes(7, 4, 256) == -4
    assert L.ggq_select_family(1, 4096, 1, 4096) == _ext.FAMILY_DECODE
    assert L.ggq_select_family(0, 16, 1, 32) == _ext.FAMILY_GENERIC      # 34-byte rows: unaligned
    assert L.ggq_select_family(2, 16, 4, 256) == _ext.FAMILY_GENERIC     # 210-byte rows
    # skinny (tcgen05, weights in TMEM): every 17 <= T <= 127, and the big T <= 16 shapes where it measured faster
    assert L.ggq_select_family(1, 4096, 40, 4096) == _ext.FAMILY_SKINNY
    assert L.ggq_select_family(1, 4096, 127, 4096) == _ext.FAMILY_SKINNY
    assert L.ggq_select_family(1, 4096, 128, 4096) == _ext.FAMILY_PREFILL
    assert L.ggq_select_family(1, 128256, 16, 4096) == _ext.FAMILY_SKINNY
    assert L.ggq_select_family(1, 128256, 8, 4096) == _ext.FAMILY_DECODE
    assert L.ggq_select_family(0, 28672, 8, 8192) == _ext.FAMILY_SKINNY
    assert L.ggq_select_family(2, 128256, 16, 4096) == _ext.FAMILY_SKINNY
    assert L.ggq_select_family(2, 128256, 8, 4096) == _ext.FAMILY_DECODE
    assert L.ggq_select_family(1, 4096, 16, 4096) == _ext.FAMILY_DECODE   # small layer: lower fixed cost
    assert b"shape" in L.ggq_error_string(-1)
    assert L.ggq_launch_count() == 0


def test_argument_errors_are_reported_without_a_device():
    from kernels import _ext
    L = _ext.lib()
    one = ctypes.c_void_p(256)
    assert L.ggq_mm_q8_0_f16(one, one, one, 4, 1, 33, None) == -1       # K % 32 (mmq_q8_0.py:124)
    assert L.ggq_mm_q4_k_f16(one, one, one, 4, 1, 128, None) == -1      # K % 256 (mmq_q4_k.py:263)
    assert L.ggq_mm_q6_k_f16(one, one, one, 4, 1, 255, None) == -1      # K % 256 (mmq_q6_k.py:211)
    assert L.ggq_mm_q4_k_f16(None, one, one, 4, 1, 256, None) == -2
    assert L.ggq_mm_q4_k_f16(None, None, None, 0, 1, 256, None) == 0    # empty problem: nothing to do
    outs = (ctypes.c_void_p * 1)(256)
    assert L.ggq_mm_ex(1, one, one, 256, outs, 1, 4, 4, 1, 256, 9, None) == -3
    assert L.ggq_mm_ex(1, one, one, 128, outs, 1, 4, 4, 1, 256, 0, None) == -1  # ldx < K


@pytest.mark.parametrize("mod,fn,consts", [
    ("kernels.mmq_q8_0", "mmq_q8_0", dict(QK8_0=32, QK8_1=32, Q8_0_SIZE=34)),
    ("kernels.mmq_q4_k", "mmq_q4_k", dict(Q4_K_BLOCK_SIZE=144, Q8_1_BLOCK_SIZE=36, Q4_K_SUBBLK_NUM=8, QK_K=256, QK8_1=32)),
    ("kernels.mmq_q6_k", "mmq_q6_k", dict(QK_K=256, Q6_K_SUBBLK_NUM=16, QK8_1=32, Q6_K_BLOCK_SIZE=210, Q8_1_BLOCK_SIZE=36)),
])
def test_entry_points_mirror_the_reference(mod, fn, consts):
    import importlib
    import inspect
    m = importlib.import_module(mod)
    f = getattr(m, fn)
    assert list(inspect.signature(f).parameters) == ["A", "B", "M", "N", "K"]
    for k, v in consts.items():
        assert getattr(m, k) == v
    A = torch.zeros(144, dtype=torch.int8)
    B = torch.zeros((1, 100), dtype=torch.float16)
    with pytest.raises(AssertionError):          # the reference's only check: K % block
        f(A, B, 1, 1, 100)
    K = 256
    with pytest.raises(ValueError):              # CPU tensors: there is no CPU path
        f(torch.zeros(1 * (K // consts.get("QK8_0", 256)) * (34 if "Q8_0_SIZE" in consts else 1), dtype=torch.int8),
          torch.zeros((1, K), dtype=torch.float16), 1, 1, K)


def test_decode_planner_invariants():
    """ggq_decode_plan is pure host code: every BASELINE decode shape gets a plan that fits one SM's shared memory;
    shapes whose activations do not fit one CTA are split across a cluster (slices = CTAs per cluster <= 8)."""
    from kernels import _ext
    L = _ext.lib()
    L.ggq_decode_plan.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.POINTER(ctypes.c_int)]
    L.ggq_decode_plan.restype = ctypes.c_int
    shapes = [(f, o, t, k) for f in (0, 1, 2) for (o, k) in ((4096, 4096), (14336, 4096), (128256, 4096), (4096, 14336),
                                                             (28672, 8192), (8192, 28672)) for t in (1, 2, 7, 8, 9, 16)]
    for f, o, t, k in shapes:
        out = (ctypes.c_int * 9)()
        assert L.ggq_decode_plan(f, o, t, k, out) == 0, (f, o, t, k)
        kw, at, nt, slices, cps, stages, grid_occ, batches, smem = list(out)
        assert smem <= 227 * 1024 and stages >= 2 and nt == (2 if t > 8 else 1), (f, o, t, k, list(out))
        assert 1 <= slices <= 8 or at == 4, (f, o, t, k, list(out))
        qk, cb = ((32, 16), (256, 2), (256, 2))[f]
        wide = f == 1 and t == 1 and k <= 14336                  # Q4_K GEMV: 4-block chunks (KGeo<1, true>) whenever they fit
        assert slices * cps * (4 if wide else cb) * qk >= k      # the slices cover K
        if wide:
            assert cps == (k // 256 + 3) // 4 and stages == 2, (f, o, t, k, list(out))
        if t <= 8 and k <= 4096:
            assert slices == 1 and at == 1, (f, o, t, k, list(out))   # every T <= 8 shape of BASELINE configs[1] is unsliced
    out = (ctypes.c_int * 9)()
    assert L.ggq_decode_plan(1, 128256, 1, 4096, out) == 0 and out[5] >= 2 and out[4] == 4   # headline: wide chunks (4 per row), >= 2 stages


def test_swiglu_and_host_pipe_argument_errors():
    """ggq_mm_swiglu / ggq_mm_host validate before anything is enqueued; the workspace query is pure host code."""
    from kernels import _ext
    L = _ext.lib()
    I64, P, INT = ctypes.c_int64, ctypes.c_void_p, ctypes.c_int
    L.ggq_mm_swiglu_workspace.argtypes = [INT, I64, I64, I64]
    L.ggq_mm_swiglu_workspace.restype = I64
    L.ggq_mm_swiglu.argtypes = [INT, P, P, P, P, I64, I64, I64, P, I64, P]
    L.ggq_mm_swiglu.restype = INT
    # decode-sized token counts on 16-byte-aligned rows run the fused kernel: no workspace
    for f, o, k in ((0, 14336, 4096), (1, 14336, 4096), (2, 14336, 4096), (1, 28672, 8192), (1, 1004, 2048)):
        for t in (1, 5, 8, 16):
            assert L.ggq_mm_swiglu_workspace(f, o, t, k) == 0, (f, o, t, k)
    assert L.ggq_mm_swiglu_workspace(1, 14336, 17, 4096) == 17 * 14336 * 2     # composed form: the gate projection
    assert L.ggq_mm_swiglu_workspace(1, 14336, 512, 4096) == 512 * 14336 * 2
    assert L.ggq_mm_swiglu_workspace(1, 14336, 1, 100) == -1 and L.ggq_mm_swiglu_workspace(5, 1, 1, 256) == -4
    one = ctypes.c_void_p(256)
    assert L.ggq_mm_swiglu(1, one, one, one, one, 64, 1, 100, None, 0, None) == -1     # K % 256
    assert L.ggq_mm_swiglu(1, one, None, one, one, 64, 1, 256, None, 0, None) == -2    # no up matrix
    assert L.ggq_mm_swiglu(1, None, None, None, None, 0, 1, 256, None, 0, None) == 0   # empty problem
    assert L.ggq_mm_swiglu(1, one, one, one, one, 64, 64, 256, None, 0, None) == -2    # composed form without a workspace
    from kernels import host
    H = host._lib()
    h = ctypes.c_void_p()
    assert H.ggq_host_pipe_create(ctypes.byref(h), 0, 16, 2) == -1
    assert H.ggq_host_pipe_create(ctypes.byref(h), 16, 16, 9) == -1
    assert H.ggq_host_pipe_create(None, 16, 16, 2) == -2
    assert H.ggq_mm_host(None, 1, one, one, one, 4, 1, 256) == -2
    assert H.ggq_host_pipe_sync(None) == -2


def test_torch_extension_loads_and_checks_operands():
    """The PyTorch extension is the binding of the per-step calls; its operand checks raise the reference-style errors."""
    from kernels import _ext
    from kernels.swiglu import mmq_q4_k_swiglu
    t = _ext.torch_ext()
    assert t.version() == _ext.lib().ggq_version()
    A = torch.zeros(144, dtype=torch.int8)
    B = torch.zeros((1, 256), dtype=torch.float16)
    with pytest.raises(ValueError):
        t.mm(1, A, B, 1, 1, 256)                      # CPU tensors
    with pytest.raises(TypeError):
        t.mm(1, A.float(), B, 1, 1, 256)
    with pytest.raises(ValueError):
        mmq_q4_k_swiglu(A, A, B, 1, 1, 256)
    with pytest.raises(AssertionError):
        mmq_q4_k_swiglu(A, A, B, 1, 1, 100)


def test_dequant_rejects_rows_beyond_32_bit_indexing():
    """K beyond what the kernels index with 32-bit arithmetic is an error, not a silent overflow."""
    from kernels import _ext
    L = _ext.lib()
    one = ctypes.c_void_p(256)
    assert L.ggq_dequant_q4_k_f16(one, one, 1, (1 << 31), None) == -1
    assert L.ggq_dequant_q8_0_f16(one, one, 1, (1 << 30) + 32, None) == -1


def test_describe_and_push_columns_host_side():
    """ggq_describe names the kernel AUTO dispatch would launch (bench.py `roofline.kernel`); ggq_push_columns validates its
    arguments before touching the GPU."""
    from kernels import _ext
    L = _ext.lib()
    I64, P, INT = ctypes.c_int64, ctypes.c_void_p, ctypes.c_int
    L.ggq_describe.argtypes = [INT, I64, I64, I64, ctypes.c_char_p, INT]
    L.ggq_describe.restype = INT
    buf = ctypes.create_string_buffer(256)

    def d(f, o, t, k):
        assert L.ggq_describe(f, o, t, k, buf, 256) == 0
        return buf.value.decode()
    assert "decode_kernel<Q4_K" in d(1, 128256, 1, 4096) and "WIDE" in d(1, 128256, 1, 4096)   # the headline kernel
    assert "WIDE" not in d(1, 128256, 8, 4096) and "WIDE" not in d(2, 128256, 1, 4096)          # Q4_K single token only
    assert "WIDE" not in d(1, 8192, 1, 28672)                    # activations do not fit next to the wide rings
    assert "skinny_kernel<Q8_0,N=16" in d(0, 28672, 16, 8192)
    assert "prefill2_kernel<Q6_K>" in d(2, 128256, 2048, 4096)
    assert "generic_kernel" in d(0, 8, 1, 32)
    assert L.ggq_describe(1, 4096, 1, 100, buf, 256) == -1 and L.ggq_describe(9, 4096, 1, 256, buf, 256) == -4
    L.ggq_push_columns.argtypes = [P, ctypes.POINTER(P), INT, I64, I64, I64, P]
    L.ggq_push_columns.restype = INT
    one = P(256)
    dst = (P * 2)(256, 512)
    assert L.ggq_push_columns(one, dst, 0, 64, 32, 4, None) == 0          # no peers: nothing to do
    assert L.ggq_push_columns(one, dst, 2, 16, 32, 4, None) == -1         # width beyond the pitch
    assert L.ggq_push_columns(one, dst, 9, 64, 32, 4, None) == -1         # more than 8 peers
    assert L.ggq_push_columns(None, dst, 2, 64, 32, 4, None) == -2
    bad = (P * 2)(256, None)
    assert L.ggq_push_columns(one, bad, 2, 64, 32, 4, None) == -2


def test_reference_arithmetic_mode_argument_errors_and_no_cpu_path():
    """ggq_mm_ref_q8_1 validates format / shape / pointers on the host; the Python mirror of kernels/cpu_impls refuses CPU
    tensors instead of falling back."""
    from kernels import _ext, q8_1_mode
    L = _ext.lib()
    L.ggq_mm_ref_q8_1.argtypes = [ctypes.c_int] + [ctypes.c_void_p] * 3 + [ctypes.c_int64] * 3 + [ctypes.c_void_p]
    L.ggq_mm_ref_q8_1.restype = ctypes.c_int
    one = ctypes.c_void_p(256)
    assert L.ggq_mm_ref_q8_1(7, one, one, one, 4, 1, 256, None) == -4      # unknown format
    assert L.ggq_mm_ref_q8_1(1, one, one, one, 4, 1, 128, None) == -1      # K % 256
    assert L.ggq_mm_ref_q8_1(0, one, one, one, 4, 1, 33, None) == -1       # K % 32
    assert L.ggq_mm_ref_q8_1(1, None, one, one, 4, 1, 256, None) == -2     # NULL weights
    assert L.ggq_mm_ref_q8_1(1, None, None, None, 0, 1, 256, None) == 0    # empty problem
    A = torch.zeros(4 * 144, dtype=torch.int8)
    B = torch.zeros(36 * 8, dtype=torch.int8)
    with pytest.raises(ValueError, match="no CPU path"):
        q8_1_mode.mmq_q4_k_q8_1(A, B, 4, 1, 256)
