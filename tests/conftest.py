import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "gguf-triton-kernel_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    import numpy as np

    def load(fmt):
        z = np.load(os.path.join(ROOT, "tests", "golden", f"{fmt}.npz"))
        n = len([k for k in z.files if k.endswith("_mnk")])
        cases = []
        for i in range(n):
            p = f"c{i}_"
            M, N, K = (int(v) for v in z[p + "mnk"])
            cases.append(dict(M=M, N=N, K=K, **{k: z[p + k] for k in "WXABDC"}))
        return cases

    return {fmt: load(fmt) for fmt in ("q8_0", "q4_k", "q6_k")}
