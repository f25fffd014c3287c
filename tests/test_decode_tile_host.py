"""Runs the decode family's lane-level device code (csrc/decode_tile.cuh) on the CPU under a 32-thread
warp emulator (tests/host) against a scalar restatement of the block formats.  Catches unpack /
fragment-layout / bias-cancellation mistakes without a GPU."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    if shutil.which("g++") is None:
        pytest.skip("g++ not available")
    exe = str(tmp_path_factory.mktemp("emu") / "emu_decode")
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-pthread", "-o", exe,
                           os.path.join(ROOT, "tests", "host", "emu_decode.cpp")])
    return exe


@pytest.mark.parametrize("fmt", [0, 1, 2])
@pytest.mark.parametrize("T,K", [(1, 2048), (8, 4096), (3, 2560 if True else 0), (16, 2048), (11, 6144)])
def test_lane_code_matches_scalar_formats(emu, fmt, T, K):
    if fmt == 0 and K % 256:
        K = 768
    if fmt == 2 and K == 2560:
        K = 4096  # Q6_K rows must be whole 16-byte vectors (K % 2048 == 0) for this family
    r = subprocess.run([emu, str(fmt), str(T), str(K), "7"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr


@pytest.mark.parametrize("fmt,K", [(1, 256), (1, 2048), (1, 2560), (1, 4096), (0, 256), (0, 768), (0, 4096), (2, 2048), (2, 4096)])
def test_single_token_gemv_lane_code(emu, fmt, K):
    """The T == 1 tile code (sub-blocks spread over the MMA columns) against the scalar format restatement."""
    r = subprocess.run([emu, str(fmt), "1", str(K), "11", "1"], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stdout + r.stderr
