"""Dev (GPU): time the reference-arithmetic (Q8_1) mode on one shape, two weight copies alternating: FMT/O/K from env, T list in argv."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gguf-triton-kernel_b200"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import torch
from kernels import q8_1_mode
from utils.quantize.q8_1 import quantize_to_q8_1
from dev_skinny import gen_weights, BLK
fmt = os.environ.get("FMT", "q4_k"); o = int(os.environ.get("O", 128256)); k = int(os.environ.get("K", 4096))
fn = {"q8_0": q8_1_mode.mmq_q8_0_q8_1, "q4_k": q8_1_mode.mmq_q4_k_q8_1, "q6_k": q8_1_mode.mmq_q6_k_q8_1}[fmt]
Ws = [gen_weights(fmt, o, k, 1 + i).view(torch.int8) for i in range(2)]
nbytes = o * (k // BLK[fmt][0]) * BLK[fmt][1]
for t in [int(a) for a in sys.argv[1:]] or [1, 8]:
    XQ = quantize_to_q8_1(torch.randn((t, k), device="cuda", dtype=torch.float16))
    for i in range(4):
        fn(Ws[i & 1], XQ, o, t, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for i in range(n):
        fn(Ws[i & 1], XQ, o, t, k)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    print(f"q8_1-mode {fmt} O={o} K={k} T={t}: {ms*1e3:.2f} us  {nbytes/ms/1e6:.0f} GB/s", flush=True)
