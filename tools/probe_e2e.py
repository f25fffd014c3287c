"""Dev (GPU): variants of the end-to-end step (pinned H2D of X, mmq, D2H of C) — which ones keep the kernels back to back."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gguf-triton-kernel_b200"), os.path.join(ROOT, "tools")):
    sys.path.insert(0, p)
import torch
from kernels import _ext as ext
from kernels.mmq_q4_k import mmq_q4_k
from dev_skinny import gen_weights
O, K, T = 128256, 4096, 1
f = ext.FMT_ID["q4_k"]
W = [gen_weights("q4_k", O, K, 1), gen_weights("q4_k", O, K, 2)]
xh = [torch.randn((T, K), dtype=torch.float16).pin_memory() for _ in range(2)]
ch = [torch.empty((T, O), dtype=torch.float16).pin_memory() for _ in range(2)]
xd = [h.cuda() for h in xh]
cd = [torch.empty((T, O), device="cuda", dtype=torch.float16) for _ in range(2)]
comp = torch.cuda.current_stream()
copy = torch.cuda.Stream()
copy2 = torch.cuda.Stream()

def run(name, step, n=200, fin=None):
    for i in range(20): step(i)
    if fin: fin()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for i in range(n): step(i)
    if fin: fin()
    e1.record()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"{name:60s} {e0.elapsed_time(e1)/n*1e3:8.2f} us/step (host issue {1e6*(t1-t0)/n:6.1f} us/step)", flush=True)

def k_only(i):
    ext.mm(f, W[i & 1], xd[0], O, T, K, out=cd[0])
run("kernel only", k_only)

ev = torch.cuda.Event(); ev.record(); 
def k_evwait(i):
    comp.wait_event(ev)
    ext.mm(f, W[i & 1], xd[0], O, T, K, out=cd[0])
run("kernel + wait_event(already done) in front", k_evwait)

def k_evrec(i):
    ext.mm(f, W[i & 1], xd[0], O, T, K, out=cd[0])
    ev.record(comp)
run("kernel + event record after", k_evrec)

def serial(i):
    xd[0].copy_(xh[0], non_blocking=True)
    ext.mm(f, W[i & 1], xd[0], O, T, K, out=cd[0])
    ch[0].copy_(cd[0], non_blocking=True)
run("A serial one stream", serial)

evx = [torch.cuda.Event() for _ in range(2)]; evk = [torch.cuda.Event() for _ in range(2)]; evc = [torch.cuda.Event() for _ in range(2)]
def two_stream(i):
    b = i & 1
    with torch.cuda.stream(copy):
        copy.wait_event(evk[b])
        xd[b].copy_(xh[b], non_blocking=True)
        evx[b].record(copy)
    comp.wait_event(evx[b]); comp.wait_event(evc[b])
    ext.mm(f, W[b], xd[b], O, T, K, out=cd[b])
    evk[b].record(comp)
    with torch.cuda.stream(copy):
        copy.wait_event(evk[b])
        ch[b].copy_(cd[b], non_blocking=True)
        evc[b].record(copy)
run("B two streams, double-buffered, ext.mm(out=)", two_stream, fin=lambda: comp.wait_stream(copy))

def three_stream(i):
    b = i & 1
    with torch.cuda.stream(copy):
        copy.wait_event(evk[b])
        xd[b].copy_(xh[b], non_blocking=True)
        evx[b].record(copy)
    comp.wait_event(evx[b]); comp.wait_event(evc[b])
    ext.mm(f, W[b], xd[b], O, T, K, out=cd[b])
    evk[b].record(comp)
    with torch.cuda.stream(copy2):
        copy2.wait_event(evk[b])
        ch[b].copy_(cd[b], non_blocking=True)
        evc[b].record(copy2)
run("B3 H2D stream + D2H stream", three_stream, fin=lambda: (comp.wait_stream(copy), comp.wait_stream(copy2)))

def alloc_rs(i):
    b = i & 1
    with torch.cuda.stream(copy):
        copy.wait_event(evk[b])
        xd[b].copy_(xh[b], non_blocking=True)
        evx[b].record(copy)
    comp.wait_event(evx[b])
    C = mmq_q4_k(W[b], xd[b], O, T, K)
    evk[b].record(comp)
    C.record_stream(copy)
    with torch.cuda.stream(copy):
        copy.wait_event(evk[b])
        ch[b].copy_(C, non_blocking=True)
run("C two streams, mmq_q4_k allocating + record_stream", alloc_rs, fin=lambda: comp.wait_stream(copy))

# zero-copy: the kernel reads X from pinned host memory and writes C to pinned host memory (UVA)
try:
    def mm_nocheck(fmt, A, B, M, N, K_, out):
        L = ext.lib()
        rc = L.ggq_mm_q4_k_f16(A.data_ptr(), B.data_ptr(), out.data_ptr(), M, N, K_, torch.cuda.current_stream().cuda_stream)
        assert rc == 0, rc
    def zc(i):
        mm_nocheck(f, W[i & 1], xh[0], O, T, K, ch[0])
    run("Z zero-copy: kernel reads/writes pinned host memory", zc)
    torch.cuda.synchronize()
    xd0 = xh[0].cuda(); ext.mm(f, W[1], xd0, O, T, K, out=cd[1]); torch.cuda.synchronize()
    print("zero-copy result equal:", torch.equal(ch[0].cuda(), cd[1]))
except Exception as e:
    print("zero-copy failed:", repr(e)[:300])
