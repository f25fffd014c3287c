"""Dev (GPU): time one family on one shape: python tools/dev_time.py fmt O K T [family] (env: GGQ_SKINNY_WU / _PROBE)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gguf-triton-kernel_b200"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import torch
from kernels import _ext as ext
from dev_skinny import gen_weights, timed, BLK
fmt, O, K, T = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
fam = int(sys.argv[5]) if len(sys.argv) > 5 else ext.FAMILY_SKINNY
nbytes = O * (K // BLK[fmt][0]) * BLK[fmt][1]
copies = max(1, min(16, -(-2 * 126_000_000 // nbytes)))
Ws = [gen_weights(fmt, O, K, 7 + i) for i in range(copies)]
X = torch.randn((T, K), device="cuda", dtype=torch.float16)
C = torch.empty((T, O), device="cuda", dtype=torch.float16)
n = max(10, copies * 2)
for i in range(copies):
    ext.mm(ext.FMT_ID[fmt], Ws[i], X, O, T, K, out=C, family=fam)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for i in range(n):
        ext.mm(ext.FMT_ID[fmt], Ws[i % copies], X, O, T, K, out=C, family=fam)
ms = timed(g.replay, n)
print(f"{fmt} O={O} K={K} T={T} fam={fam} WU={os.environ.get('GGQ_SKINNY_WU','-')} probe={os.environ.get('GGQ_SKINNY_PROBE','0')}: "
      f"{ms*1e3:.2f} us  {nbytes/ms/1e6:.0f} GB/s", flush=True)
