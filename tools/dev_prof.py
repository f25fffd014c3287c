import os, sys
os.environ["GGQ_SKINNY_PROF"] = "1"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gguf-triton-kernel_b200"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import torch
from kernels import _ext as ext
from dev_skinny import gen_weights
fmt, O, K, T = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
W = gen_weights(fmt, O, K, 1)
X = torch.randn((T, K), device="cuda", dtype=torch.float16)
for _ in range(2):
    ext.mm(ext.FMT_ID[fmt], W, X, O, T, K, family=ext.FAMILY_SKINNY)
torch.cuda.synchronize()
