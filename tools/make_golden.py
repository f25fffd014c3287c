#!/usr/bin/env python
"""Mint golden vectors for the oracle from the REFERENCE ITSELF (run in the build container only).

The reference holds no golden vectors, known-answer tests or fixtures (its tests are unseeded
random differential checks, test/test_mmq_q8_0.py:27-36), so parity is pinned on outputs of the
reference code run here: this script copies /root/reference to a scratch dir (it is read-only and
its ctypes wrappers look for the packer .so next to themselves, utils/quantize/q4_k.py:39-46),
builds the two C packers there, imports the reference's Python modules and records, for seeded
inputs:

  * the packed weight bytes its packers produce          (utils/quantize/q8_0.py, q4_k.py, q6_k.py)
  * the Q8_1 activation bytes                            (utils/quantize/q8_1.py)
  * its dequantizer outputs                              (dequantize_q8_0 / q4_k / q6_k)
  * its CPU mmq outputs                                  (kernels/cpu_impls/*)

into tests/golden/<fmt>.npz.  /root/reference does not exist on the GPU box; the fixtures travel.

    python tools/make_golden.py            # writes tests/golden/{q8_0,q4_k,q6_k}.npz
"""
import os
import shutil
import subprocess
import sys
import tempfile

import numpy as np

REF = os.environ.get("GGQ_REFERENCE", "/root/reference")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# (M = out-features, N = tokens, K): the reference test grids' corners (test/test_mmq_*.py:17-21)
# plus a K large enough to exercise the fp16 accumulator and multi-block rows.
CASES = {
    "q8_0": [(1, 1, 32), (4, 1, 64), (16, 4, 128), (4, 16, 512), (16, 16, 256), (8, 3, 1024), (5, 2, 4096)],
    "q4_k": [(1, 1, 256), (4, 4, 512), (16, 1, 1024), (1, 16, 1024), (16, 16, 256), (6, 3, 2048), (3, 2, 4096)],
    "q6_k": [(1, 1, 256), (4, 4, 512), (16, 1, 1024), (1, 16, 1024), (16, 16, 256), (6, 3, 2048), (3, 2, 4096)],
}


def stage() -> str:
    tmp = tempfile.mkdtemp(prefix="ggq_ref_")
    dst = os.path.join(tmp, "ref")
    shutil.copytree(REF, dst)
    qd = os.path.join(dst, "utils", "quantize")
    for n in ("q4_k", "q6_k"):
        subprocess.check_call(["gcc", "-O2", "-shared", "-fPIC", "-o", os.path.join(qd, f"lib{n}_ref.so"),
                               os.path.join(qd, f"{n}_ref.c"), "-lm"])
    return dst


def main():
    import torch
    ref = stage()
    sys.path.insert(0, ref)
    from utils.quantize.q8_0 import quantize_to_q8_0, dequantize_q8_0
    from utils.quantize.q8_1 import quantize_to_q8_1
    from utils.quantize.q4_k import quantize_to_q4_k, dequantize_q4_k
    from utils.quantize.q6_k import quantize_to_q6_k, dequantize_q6_k
    from kernels.cpu_impls.mmq_q8_0_q8_1_cpu import mmq_q8_0_q8_1_cpu
    from kernels.cpu_impls.mmq_q4_k_q8_1_cpu import mmq_q4_k_q8_1_cpu
    from kernels.cpu_impls.mmq_q6_k_q8_1_cpu import mmq_q6_k_q8_1_cpu

    quant = {"q8_0": quantize_to_q8_0, "q4_k": quantize_to_q4_k, "q6_k": quantize_to_q6_k}
    dequant = {"q8_0": dequantize_q8_0, "q4_k": dequantize_q4_k, "q6_k": dequantize_q6_k}
    mmq = {"q8_0": mmq_q8_0_q8_1_cpu, "q4_k": mmq_q4_k_q8_1_cpu, "q6_k": mmq_q6_k_q8_1_cpu}

    os.makedirs(OUT, exist_ok=True)
    for fmt, cases in CASES.items():
        blob = {}
        for i, (M, N, K) in enumerate(cases):
            torch.manual_seed(42 + i)  # 42: the seed the oracle demos use (mmq_q4_k_q8_1_cpu.py:124)
            scale = [1.0, 0.05, 3.0][i % 3]  # vary magnitudes so fp16 scales cover several binades
            W = (torch.randn(M, K) * scale).to(torch.float16)
            X = torch.randn(N, K).to(torch.float16)
            A = quant[fmt](W)
            B = quantize_to_q8_1(X)
            D = dequant[fmt](A, (M, K))
            C = mmq[fmt](A, B, M, N, K)
            p = f"c{i}_"
            blob[p + "mnk"] = np.array([M, N, K], dtype=np.int64)
            blob[p + "W"] = W.numpy()
            blob[p + "X"] = X.numpy()
            blob[p + "A"] = A.numpy().astype(np.int8)
            blob[p + "B"] = B.numpy().astype(np.int8)
            blob[p + "D"] = D.numpy()
            blob[p + "C"] = C.contiguous().numpy()
            print(fmt, (M, N, K), "A", A.numel(), "C", tuple(C.shape), D.dtype)
        np.savez_compressed(os.path.join(OUT, f"{fmt}.npz"), **blob)
    shutil.rmtree(os.path.dirname(ref))


if __name__ == "__main__":
    main()
