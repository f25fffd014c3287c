import sys, os, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/gguf-triton-kernel_b200")
import bench
from kernels import _ext as ext
for fmt,o,k in (("q4_k",128256,4096),("q6_k",128256,4096),("q8_0",28672,8192)):
    W = bench.gen_weights(torch, fmt, o, k, "cuda", 1)
    for t in (1,):
        X = torch.randn((t,k), device="cuda", dtype=torch.float16); C = torch.empty((t,o), device="cuda", dtype=torch.float16)
        ms = bench.timed(torch, None, lambda: ext.mm(ext.FMT_ID[fmt], W, X, o, t, k, out=C), 30, 5, 1)
        nb = bench.packed_bytes(fmt,o,k)
        print(os.environ.get("GGQ_DECODE_NOCOMPUTE","0"), fmt, o, k, t, round(ms*1e3,1), "us", round(nb/ms/1e6,1), "GB/s")
