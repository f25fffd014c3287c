"""The K-quant packer arithmetic (csrc/kquant_pack.cuh — what the GPU packers run one thread per sub-block), compiled
for the host, against the reference's compiled packers (oracle/_ref, built from /root/reference by `make -C oracle
ref`): byte-identical on every distribution below.  The GPU kernels themselves are compared in
tests/test_gpu_quantize_ops.py."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")


@pytest.fixture(scope="module")
def libs(tmp_path_factory):
    if not os.path.exists(os.path.join(REF, "libq4_k_ref.so")):
        pytest.skip("oracle/_ref not built (needs /root/reference: make -C oracle ref)")
    so = str(tmp_path_factory.mktemp("kq") / "libkquant_host.so")
    subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-o", so,
                           os.path.join(ROOT, "tests", "host", "kquant_host.cpp")])
    return ctypes.CDLL(so), ctypes.CDLL(os.path.join(REF, "libq4_k_ref.so")), ctypes.CDLL(os.path.join(REF, "libq6_k_ref.so"))


def _run(lib, fn, x, blk):
    x = np.ascontiguousarray(x, dtype=np.float32)
    out = np.zeros(x.size // 256 * blk, dtype=np.uint8)
    getattr(lib, fn)(x.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p), ctypes.c_longlong(x.size))
    return out


def _cases():
    rng = np.random.default_rng(0)
    n = 256 * 256
    return {
        "randn": rng.standard_normal(n).astype(np.float32),
        "randn fp16-valued": rng.standard_normal(n).astype(np.float16).astype(np.float32),
        "uniform positive": rng.uniform(0, 3, n).astype(np.float32),
        "uniform negative": rng.uniform(-3, 0, n).astype(np.float32),
        "tiny": (rng.standard_normal(n) * 1e-20).astype(np.float32),
        "large": (rng.standard_normal(n) * 3e4).astype(np.float32),
        "zeros": np.zeros(1024, dtype=np.float32),
        "constant": np.full(1024, 0.37, dtype=np.float32),
        "negative constant": np.full(1024, -1.5, dtype=np.float32),
        "sparse": np.where(rng.random(n) < 0.1, rng.standard_normal(n), 0).astype(np.float32),
        "heavy tail": np.clip(rng.standard_t(2, n), -1e4, 1e4).astype(np.float32),
        "mixed scales": (rng.standard_normal(n) * np.repeat(10.0 ** rng.uniform(-6, 3, n // 32), 32)).astype(np.float32),
        "one hot": (np.eye(256, dtype=np.float32) * 2.5).ravel(),
    }


@pytest.mark.parametrize("name", list(_cases()))
def test_host_build_is_byte_identical_to_the_reference_packers(libs, name):
    mine, ref4, ref6 = libs
    x = _cases()[name]
    assert np.array_equal(_run(mine, "host_quantize_q4_k", x, 144), _run(ref4, "quantize_row_q4_K_ref", x, 144))
    assert np.array_equal(_run(mine, "host_quantize_q6_k", x, 210), _run(ref6, "quantize_row_q6_K_ref", x, 210))
