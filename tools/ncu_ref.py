"""dev helper: launch the reference-arithmetic (Q8_1) mode a few times (target of an ncu capture). FMT/O/K/T from env."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gguf-triton-kernel_b200"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import torch
from kernels import q8_1_mode
from utils.quantize.q8_1 import quantize_to_q8_1
from dev_skinny import gen_weights
fmt = os.environ.get("FMT", "q4_k"); o = int(os.environ.get("O", 128256)); k = int(os.environ.get("K", 4096)); t = int(os.environ.get("T", 1))
fn = {"q8_0": q8_1_mode.mmq_q8_0_q8_1, "q4_k": q8_1_mode.mmq_q4_k_q8_1, "q6_k": q8_1_mode.mmq_q6_k_q8_1}[fmt]
W = gen_weights(fmt, o, k, 1).view(torch.int8)
XQ = quantize_to_q8_1(torch.randn((t, k), device="cuda", dtype=torch.float16))
for _ in range(4):
    fn(W, XQ, o, t, k)
torch.cuda.synchronize()
print("done")
