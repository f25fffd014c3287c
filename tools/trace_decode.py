"""Dev (GPU): per-CTA phase timeline of back-to-back decode launches (ggq_dev_set_trace).
python tools/trace_decode.py fmt O K T [n_launches]"""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gguf-triton-kernel_b200"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import numpy as np
import torch
from kernels import _ext as ext
from dev_skinny import gen_weights
fmt, O, K, T = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
n = int(sys.argv[5]) if len(sys.argv) > 5 else 6
L = ext.lib()
L.ggq_dev_set_trace.argtypes = [ctypes.c_void_p]
L.ggq_dev_set_trace.restype = None
copies = 4
Ws = [gen_weights(fmt, O, K, i) for i in range(copies)]
X = torch.randn((T, K), device="cuda", dtype=torch.float16)
C = torch.empty((T, O), device="cuda", dtype=torch.float16)
f = ext.FMT_ID[fmt]
for i in range(copies):
    ext.mm(f, Ws[i], X, O, T, K, out=C, family=ext.FAMILY_DECODE)
torch.cuda.synchronize()
buf = torch.zeros((16, 320, 8), dtype=torch.int64, device="cuda")
L.ggq_dev_set_trace(buf.data_ptr())
g = torch.cuda.CUDAGraph()
with torch.cuda.graph(g):
    for i in range(n):
        ext.mm(f, Ws[i % copies], X, O, T, K, out=C, family=ext.FAMILY_DECODE)
L.ggq_dev_set_trace(None)
g.replay(); torch.cuda.synchronize()
buf.zero_()
g.replay(); torch.cuda.synchronize()
t = buf.cpu().numpy().astype(np.int64)
names = ["entry", "boxes", "pdlwait", "xstaged", "w0done", "ctadone"]
t0 = t[0, :, 0][t[0, :, 0] > 0].min()
print(f"{fmt} O={O} K={K} T={T}: ns relative to the first CTA entry of launch 0; per launch min / median / max over CTAs")
prev_end = None
for l in range(n):
    grid = int((t[l, :, 0] > 0).sum())
    row = [f"launch {l} grid={grid}"]
    for e, nm in enumerate(names):
        v = t[l, :grid, e] - t0
        row.append(f"{nm} {v.min():6d}/{int(np.median(v)):6d}/{v.max():6d}")
    end = (t[l, :grid, 5] - t0).max()
    row.append(f"| span {end - (t[l, :grid, 0] - t0).min():6d}" + (f" period {end - prev_end:6d}" if prev_end is not None else ""))
    prev_end = end
    print("  ".join(row))
