"""One prefill GEMM per format on a 2-wave problem (for ncu)."""
import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gguf-triton-kernel_b200"))
import bench
from kernels import _ext as ext
fmt = sys.argv[1] if len(sys.argv) > 1 else "q4_k"
o, k, t = 9472, 4096, 2048
W = bench.gen_weights(torch, fmt, o, k, "cuda", 11)
X = torch.randn((t, k), device="cuda", dtype=torch.float16)
C = torch.empty((t, o), device="cuda", dtype=torch.float16)
for _ in range(3):
    ext.mm(ext.FMT_ID[fmt], W, X, o, t, k, out=C)
torch.cuda.synchronize()
ms = bench.timed(torch, None, lambda: ext.mm(ext.FMT_ID[fmt], W, X, o, t, k, out=C), 5, 3, 1)
print(fmt, os.environ.get("GGQ_PREFILL_1CTA", "0"), ms, 2.0 * t * o * k / ms / 1e9, "TFLOPs")
