"""dev helper: launch the fused SwiGLU up-projection a few times (target of an ncu capture). FMT/O/K/T from env."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gguf-triton-kernel_b200"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import torch
from kernels.swiglu import mmq_swiglu
from dev_skinny import gen_weights
fmt = os.environ.get("FMT", "q4_k"); o = int(os.environ.get("O", 28672)); k = int(os.environ.get("K", 8192)); t = int(os.environ.get("T", 1))
Wg, Wu = gen_weights(fmt, o, k, 1), gen_weights(fmt, o, k, 2)
X = torch.randn((t, k), device="cuda", dtype=torch.float16) * 0.05
for _ in range(6):
    mmq_swiglu(fmt, Wg, Wu, X, o, t, k)
torch.cuda.synchronize()
print("done")
