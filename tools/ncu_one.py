"""dev helper: launch one decode shape a few times (target of an ncu capture). FMT/O/K/T from env."""
import sys, os, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/gguf-triton-kernel_b200")
import bench
from kernels import _ext as ext
v = os.environ.get("GGQ_VARIANT", "")
if v:
    ext._LIB_PATH = f"/root/repo/build_variants/libggq_{v}.so"
fmt = os.environ.get("FMT", "q4_k"); o = int(os.environ.get("O", 128256)); k = int(os.environ.get("K", 4096)); t = int(os.environ.get("T", 1))
W = bench.gen_weights(torch, fmt, o, k, "cuda", 1)
X = torch.randn((t, k), device="cuda", dtype=torch.float16); C = torch.empty((t, o), device="cuda", dtype=torch.float16)
for _ in range(6):
    ext.mm(ext.FMT_ID[fmt], W, X, o, t, k, out=C)
torch.cuda.synchronize()
print("done")
