"""Dev (GPU): one call of a family on one shape, result checked loosely: python tools/dev_one.py fmt O K T [family]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gguf-triton-kernel_b200"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import torch
from kernels import _ext as ext
from dev_skinny import gen_weights, ref, errs
fmt, O, K, T = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
fam = int(sys.argv[5]) if len(sys.argv) > 5 else ext.FAMILY_SKINNY
W = gen_weights(fmt, O, K, 3)
X = torch.randn((T, K), device="cuda", dtype=torch.float16)
for _ in range(3):
    C = ext.mm(ext.FMT_ID[fmt], W, X, O, T, K, family=fam)
torch.cuda.synchronize()
print(fmt, O, K, T, "errs", errs(C, ref(fmt, W, X, O, K)), flush=True)
