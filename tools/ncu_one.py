"""dev helper: launch one shape a few times (target of an ncu capture). FMT/O/K/T/FAMILY from env."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gguf-triton-kernel_b200"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import torch
from kernels import _ext as ext
from dev_skinny import gen_weights
fmt = os.environ.get("FMT", "q4_k"); o = int(os.environ.get("O", 128256)); k = int(os.environ.get("K", 4096)); t = int(os.environ.get("T", 1))
fam = int(os.environ.get("FAMILY", 0))
W = gen_weights(fmt, o, k, 1)
X = torch.randn((t, k), device="cuda", dtype=torch.float16); C = torch.empty((t, o), device="cuda", dtype=torch.float16)
for _ in range(6):
    ext.mm(ext.FMT_ID[fmt], W, X, o, t, k, out=C, family=fam)
torch.cuda.synchronize()
print("done")
