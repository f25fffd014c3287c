"""dev helper: CUDA-graph-timed decode shapes (device time per launch, rotating weight copies > L2)."""
import sys, os, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/gguf-triton-kernel_b200")
import bench
from kernels import _ext as ext
v = os.environ.get("GGQ_VARIANT", "")
if v:
    ext._LIB_PATH = f"/root/repo/build_variants/libggq_{v}.so"
shapes = [("q4_k",4096,4096),("q4_k",6144,4096),("q4_k",14336,4096),("q4_k",28672,4096),("q4_k",4096,14336),("q6_k",4096,14336),("q6_k",4096,4096),("q8_0",4096,4096),("q8_0",14336,4096),("q4_k",8192,8192),("q4_k",28672,8192),("q4_k",8192,28672)]
Ts = [int(t) for t in os.environ.get("TS", "1,4").split(",")]
tag = os.environ.get("TAG", v or "base")
for fmt,o,k in shapes:
    nbytes = bench.packed_bytes(fmt,o,k)
    copies = max(1, min(32, -(-2*126_000_000 // nbytes)))
    Ws = [bench.gen_weights(torch, fmt, o, k, "cuda", 7+i) for i in range(copies)]
    for t in Ts:
        X = torch.randn((t,k), device="cuda", dtype=torch.float16); C = torch.empty((t,o), device="cuda", dtype=torch.float16)
        n = max(16, copies*2)
        for i in range(min(copies,3)): ext.mm(ext.FMT_ID[fmt], Ws[i], X, o, t, k, out=C)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(n): ext.mm(ext.FMT_ID[fmt], Ws[i % copies], X, o, t, k, out=C)
        ms = bench.timed(torch, None, g.replay, 5, 3, 1) / n
        print(tag, fmt, o, k, t, round(ms*1e3,2), "us", round(nbytes/ms/1e6,1), "GB/s", flush=True)
    del Ws; torch.cuda.empty_cache()
