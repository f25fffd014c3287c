"""Dev harness (GPU): parity + timing of the skinny tcgen05 family against the mma.sync decode family."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "gguf-triton-kernel_b200")):
    sys.path.insert(0, p)
import numpy as np
import torch

from kernels import _ext as ext

torch.backends.cuda.matmul.allow_tf32 = False
BLK = {"q8_0": (32, 34), "q4_k": (256, 144), "q6_k": (256, 210)}


def gen_weights(fmt, rows, k, seed):
    qk, blk = BLK[fmt]
    nb = rows * (k // qk)
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    raw = torch.randint(0, 256, (nb, blk), dtype=torch.uint8, device="cuda", generator=g)

    def scales(mag):
        v = (torch.rand(nb, device="cuda", generator=g) * 0.75 + 0.25) * mag
        return v.to(torch.float16).view(torch.uint8).reshape(nb, 2)

    if fmt == "q8_0":
        raw[:, 0:2] = scales(0.02)
    elif fmt == "q4_k":
        raw[:, 0:2] = scales(0.02 / 16)
        raw[:, 2:4] = scales(0.02 / 16)
    else:
        raw[:, 208:210] = scales(0.02 / 64)
    return raw.reshape(-1).view(torch.int8)


def ref(fmt, W, X, O, K, rows=None):
    f = ext.FMT_ID[fmt]
    if rows is not None:
        rb = K // BLK[fmt][0] * BLK[fmt][1]
        W = W.view(-1, rb)[rows].reshape(-1).contiguous()
        O = len(rows)
    D = ext.dequant(f, W, O, K).double()
    return X.double() @ D.T


def errs(C, R):
    C = C.double()
    mx = (C - R).abs().max().item() / max(R.abs().max().item(), 1e-30)
    fro = (C - R).norm().item() / max(R.norm().item(), 1e-30)
    return mx, fro


def parity():
    bad = 0
    shapes = [(128, 2, 256), (128, 16, 1024), (256, 16, 2048), (100, 3, 4096), (1000, 8, 2048), (129, 16, 4096),
              (4096, 16, 4096), (2500, 5, 4096), (333, 32, 2048), (777, 64, 4096), (515, 100, 2048), (300, 127, 1024),
              (16, 4, 512), (28672 // 4, 16, 8192), (128256, 2, 4096), (20000, 16, 2048), (9000, 30, 2304), (5000, 7, 256),
              (70000, 16, 256), (3000, 16, 16384)]
    for fmt in ("q4_k", "q8_0", "q6_k"):
        for (O, T, K) in shapes:
            W = gen_weights(fmt, O, K, O + T)
            X = torch.randn((T, K), device="cuda", dtype=torch.float16)
            try:
                C = ext.mm(ext.FMT_ID[fmt], W, X, O, T, K, family=ext.FAMILY_SKINNY)
                torch.cuda.synchronize()
            except Exception as e:
                if "(-3)" in repr(e):
                    print("skip", fmt, O, T, K, "(unsupported shape)")
                    continue
                print("FAIL", fmt, O, T, K, repr(e)[:200])
                bad += 1
                continue
            if O > 4096:   # big layers: reference on sampled rows (first / last tiles + random)
                g = torch.Generator(device="cuda"); g.manual_seed(1)
                rows = torch.cat([torch.arange(0, 256, device="cuda"), torch.arange(O - 256, O, device="cuda"),
                                  torch.randint(0, O, (1536,), device="cuda", generator=g)])
                mx, fro = errs(C[:, rows], ref(fmt, W, X, O, K, rows=rows))
            else:
                mx, fro = errs(C, ref(fmt, W, X, O, K))
            ok = mx <= 1e-2 and fro <= 2e-3
            # determinism / repeated launches (workspace flags must be reset)
            C2 = ext.mm(ext.FMT_ID[fmt], W, X, O, T, K, family=ext.FAMILY_SKINNY)
            C3 = ext.mm(ext.FMT_ID[fmt], W, X, O, T, K, family=ext.FAMILY_SKINNY)
            torch.cuda.synchronize()
            same = torch.equal(C, C2) and torch.equal(C, C3)
            print(("ok  " if ok and same else "BAD "), fmt, O, T, K, f"max={mx:.2e} fro={fro:.2e} repeat_equal={same}", flush=True)
            bad += not (ok and same)
    return bad


def timed(fn, n_inner, reps=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    best = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best.append(e0.elapsed_time(e1) / n_inner)
    return sorted(best)[len(best) // 2]


def bench(shapes, Ts, families):
    out = []
    for fmt, O, K in shapes:
        nbytes = O * (K // BLK[fmt][0]) * BLK[fmt][1]
        copies = max(1, min(64, -(-2 * 126_000_000 // nbytes)))
        Ws = [gen_weights(fmt, O, K, 7 + i) for i in range(copies)]
        for T in Ts:
            X = torch.randn((T, K), device="cuda", dtype=torch.float16)
            C = torch.empty((T, O), device="cuda", dtype=torch.float16)
            row = {"fmt": fmt, "O": O, "K": K, "T": T}
            for name, fam in families:
                n = max(16, copies * 2)
                try:
                    for i in range(min(copies, 3)):
                        ext.mm(ext.FMT_ID[fmt], Ws[i], X, O, T, K, out=C, family=fam)
                    torch.cuda.synchronize()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        for i in range(n):
                            ext.mm(ext.FMT_ID[fmt], Ws[i % copies], X, O, T, K, out=C, family=fam)
                    ms = timed(g.replay, n)
                    row[name + "_us"] = round(ms * 1e3, 2)
                    row[name + "_GBps"] = round(nbytes / ms / 1e6, 1)
                except Exception as e:
                    row[name + "_err"] = repr(e)[:120]
            print(json.dumps(row), flush=True)
            out.append(row)
        del Ws
        torch.cuda.empty_cache()
    return out


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    bad = 0
    if what in ("all", "parity"):
        bad = parity()
        print("parity failures:", bad, flush=True)
    if what in ("all", "bench") and bad == 0:
        big = [("q4_k", 128256, 4096), ("q6_k", 128256, 4096), ("q8_0", 28672, 8192)]
        small = [("q8_0", 4096, 4096), ("q4_k", 4096, 4096), ("q4_k", 14336, 4096), ("q6_k", 4096, 14336)]
        fams = [("skinny", ext.FAMILY_SKINNY), ("decode", ext.FAMILY_DECODE)]
        bench(big, (2, 8, 16), fams)
        bench(big, (32, 64, 127), [("skinny", ext.FAMILY_SKINNY)])
        bench(small, (8, 16), fams)
    sys.exit(1 if bad else 0)
