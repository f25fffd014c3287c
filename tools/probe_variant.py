"""dev helper: time decode shapes with a chosen build (GGQ_VARIANT=name -> build_variants/libggq_<name>.so)."""
import sys, os, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/gguf-triton-kernel_b200")
import bench
from kernels import _ext as ext
v = os.environ.get("GGQ_VARIANT", "")
if v:
    ext._LIB_PATH = f"/root/repo/build_variants/libggq_{v}.so"
shapes = [("q4_k",128256,4096,1),("q4_k",128256,4096,4),("q4_k",128256,4096,8),("q4_k",128256,4096,16),("q4_k",4096,4096,1),("q4_k",14336,4096,1),("q4_k",4096,14336,1),
          ("q6_k",128256,4096,1),("q8_0",28672,8192,1)]
if os.environ.get("GGQ_SHAPES") == "q4":
    shapes = [s for s in shapes if s[0] == "q4_k"]
for fmt,o,k,t in shapes:
    W = bench.gen_weights(torch, fmt, o, k, "cuda", 1)
    X = torch.randn((t,k), device="cuda", dtype=torch.float16); C = torch.empty((t,o), device="cuda", dtype=torch.float16)
    ext.mm(ext.FMT_ID[fmt], W, X, o, t, k, out=C)
    ref = (X.float() @ ext.dequant(ext.FMT_ID[fmt], W[: 64 * (k // ext.FMT_QK[ext.FMT_ID[fmt]]) * ext.FMT_BLK[ext.FMT_ID[fmt]]], 64, k).float().t())
    err = ((C[:, :64].float() - ref).norm() / ref.norm()).item()
    ms = bench.timed(torch, None, lambda: ext.mm(ext.FMT_ID[fmt], W, X, o, t, k, out=C), 30, 5, 1)
    nb = bench.packed_bytes(fmt,o,k)
    print(v or "base", fmt, o, k, t, round(ms*1e3,1), "us", round(nb/ms/1e6,1), "GB/s", "relerr %.1e" % err, flush=True)
