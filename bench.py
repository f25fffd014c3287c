#!/usr/bin/env python
"""bench.py — benchmark of the GGUF mmq hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--no-cells] [--detail]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): decode GEMV HBM GB/s over packed weight bytes (headline) and, in `cells`, every cell
of "decode GB/s & prefill TFLOPS per Q8_0/Q4_K/Q6_K".

Headline workload at every N: BASELINE configs[1], the Llama-3-8B lm_head shape — Q4_K weights
[O=128256, K=4096] (295.5 MB packed) times T=1 fp16 activations.  The weights are synthetic fp32 values
packed by the repo's own GPU packers (utils/quantize, byte-identical to the reference's).  A step = one pass
of the hot path over one batch of synthetic activations:
  N = 1  one `mmq_q4_k` call; consecutive steps alternate between two weight copies (591 MB > 126 MB L2)
  N > 1  the layer is N-split (each rank holds O/N packed rows): activations pushed from rank 0, per-rank
         GEMV on the shard, [T, O/N] slices exchanged inside the kernel over NVLink -> "scaling": "strong".
         Consecutive steps rotate through enough shard copies that a shard never stays in L2.
`value` = packed bytes of the whole layer / step time (device-timed, inputs resident in HBM, max over ranks).
`e2e` = the same through the reference-named Python entry point with HOST activations and a HOST result
(pinned H2D of X and D2H of C inside the timed region, on a copy stream, double-buffered; the packed weights
are the layer's resident state, as in the reference's own usage).
`cells` = one entry per metric cell (decode: Q8_0 FFN / Q4_K lm_head / Q6_K lm_head at T = 1, 8, 16; prefill:
Q4_K + Q8_0 at T=4096, K=8192, O=28672 and Q6_K at T=2048; config 1; config 5's layers), each timed on the
device and checked in this run against the fp32 reference on sampled rows (oracle = checker only).
`--impl reference` times the reference's CPU path (the oracle port of kernels/cpu_impls, all host threads)
on bounded row samples of the same workload.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "gguf-triton-kernel_b200")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

FMT, O, K, T = "q4_k", 128256, 4096, 1
WORKLOAD = "Q4_K decode GEMV, Llama-3-8B lm_head (K=4096, O=128256), T=1 [BASELINE configs[1]]"
METRIC = "decode GEMV HBM GB/s over packed weight bytes (Q4_K, T=1)"
BLK = {"q8_0": (32, 34), "q4_k": (256, 144), "q6_k": (256, 210)}
L2_BYTES = 126_000_000


def packed_bytes(fmt, rows, k):
    qk, blk = BLK[fmt]
    return rows * (k // qk) * blk


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            m = json.load(f)
        return float(m["hbm_gbs"]), float(m.get("bf16_tflops", 1590.0)), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, 1590.0, "fallback (B200_PROFILING.md)"


def traffic_sidecar():
    """dram bytes per launch from the committed ncu --set full captures (tools/ncu_traffic.py writes the file)."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return {}


# ---------------------------------------------------------------------------------------------
# synthetic weights: random values packed by the repo's GPU packers (what BASELINE's configs state)
# ---------------------------------------------------------------------------------------------
def make_weights(torch, fmt, rows, k, seed, device="cuda"):
    """fp32 N(0, 0.02) weights [rows, k] -> packed flat int8, through utils/quantize (chunks of rows bound the
    temporary fp32 memory)."""
    from utils.quantize.q4_k import quantize_to_q4_k
    from utils.quantize.q6_k import quantize_to_q6_k
    from utils.quantize.q8_0 import quantize_to_q8_0
    pack = {"q8_0": quantize_to_q8_0, "q4_k": quantize_to_q4_k, "q6_k": quantize_to_q6_k}[fmt]
    g = torch.Generator(device=device)
    g.manual_seed(seed)
    out = torch.empty(packed_bytes(fmt, rows, k), dtype=torch.int8, device=device)
    rb = packed_bytes(fmt, 1, k)
    step = max(1, (1 << 28) // k)
    for r0 in range(0, rows, step):
        r1 = min(rows, r0 + step)
        w = torch.randn((r1 - r0, k), device=device, dtype=torch.float32, generator=g) * 0.02
        if fmt == "q8_0":
            w = w.to(torch.float16)
        out[r0 * rb:r1 * rb] = pack(w)
        del w
    return out


def sample_rows(rows, n, seed=0):
    """first / last rows (tile edges) plus random ones"""
    n = min(n, rows)
    rng = np.random.default_rng(seed)
    edge = [i for i in (0, 1, rows - 2, rows - 1) if 0 <= i < rows]
    pick = np.unique(np.concatenate([np.array(edge, dtype=np.int64), rng.choice(rows, n, replace=False)]))
    return pick


def ref_rows(torch, fmt, W, k, pick, x_np):
    """fp32-accumulated reference (oracle.ref32) of the sampled output rows: float32 [T, len(pick)]"""
    from oracle import ggq_oracle as orc
    rb = packed_bytes(fmt, 1, k)
    Wc = W.view(-1, rb)[torch.from_numpy(pick).to(W.device)].cpu().numpy().reshape(-1)
    return orc.ref32(fmt, Wc, x_np, len(pick), x_np.shape[0], k)


def parity_of(torch, C, pick, ref, col0=0):
    """Tier-1 errors of C[:, col0 + pick] against ref; raises when outside the north-star tolerance."""
    from oracle import ggq_oracle as orc
    got = C[:, torch.from_numpy(pick + col0).to(C.device)].float().cpu().numpy()
    mx, fro = orc.tier1_errors(got, ref)
    ok = bool(mx <= orc.TIER1_MAX and fro <= orc.TIER1_FRO)
    return {"max_over_max": float(f"{mx:.3e}"), "rel_fro": float(f"{fro:.3e}"), "rows_sampled": int(len(pick)), "ok": ok}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed regions."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: [self.lines.append(l) for l in self.proc.stdout], daemon=True).start()
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
                for n, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


# ---------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port of kernels/cpu_impls on the host cores
# ---------------------------------------------------------------------------------------------
def cpu_path_gbs(rows_per_step, steps, warmup, threads):
    """Times the reference's CPU path (Q8_1-quantized activations, integer block dots, fp16 accumulate —
    kernels/cpu_impls/mmq_q4_k_q8_1_cpu.py:61-119, restated in oracle/ggq_oracle.py) on `rows_per_step`
    rows of the workload per step, rows spread over `threads` host threads."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle import ggq_oracle as orc
    rows_per_step = max(threads, rows_per_step // threads * threads)
    per = rows_per_step // threads
    A = [orc.random_blocks(FMT, per, K, seed=100 + i) for i in range(threads)]
    X = np.random.default_rng(0).standard_normal((T, K)).astype(np.float16)

    def step():
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(lambda a: orc.mmq_cpu(FMT, a, X, per, T, K), A))

    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt = (time.perf_counter() - t0) / steps
    return packed_bytes(FMT, rows_per_step, K) / dt / 1e9, dt, rows_per_step


PORT_NOTE = ("numpy port of kernels/cpu_impls (vectorised; ~2000x faster than the reference's own pure-Python "
             "triple loop, so GPU/CPU ratios against it are conservative)")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    rows = 1024 * threads  # ~0.1 s of CPU work per step: any --steps/--warmup ends within minutes
    gbs, dt, rows = cpu_path_gbs(rows, args.steps, args.warmup, threads)
    sample = f"{rows} of {O} rows per step (rows are independent), {threads} threads; {PORT_NOTE}"
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": gbs, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "ms_full_workload_extrapolated": dt * 1e3 * O / rows,
        "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f16", "arithmetic": "Q8_1 activations, int8 block dots, fp16 accumulate (reference CPU arithmetic)",
        "data": "synthetic",
        "config": {"workload": WORKLOAD, "fmt": FMT, "O": O, "K": K, "T": T},
        "cpu_baseline": {"value": gbs, "unit": "GB/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": gbs, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ---------------------------------------------------------------------------------------------
def timed(torch, dist, fn, steps, warmup, world):
    """W untimed calls, then exactly `steps` calls between CUDA events on the current stream, barrier +
    synchronize on both sides, max over ranks.  ms per call."""
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    if world > 1:
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
        dist.barrier()
    return ms


def describe(ext, fmt, o, t, k):
    L = ext.lib()
    buf = ctypes.create_string_buffer(256)
    L.ggq_describe.argtypes = [ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_char_p, ctypes.c_int]
    L.ggq_describe.restype = ctypes.c_int
    rc = L.ggq_describe(ext.FMT_ID[fmt], o, t, k, buf, 256)
    return buf.value.decode() if rc == 0 else f"ggq_describe failed ({rc})"


def graph_of(torch, calls):
    """One CUDA graph of the given launches (removes the Python/ctypes launch cost from device-side numbers)."""
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for c in calls:
            c()
    return g


def decode_cell(torch, ext, name, fmt, o, k, ts, hbm_peak, W=None, seed=7):
    """Decode GB/s of one layer for each T in `ts`: a CUDA graph of launches cycling through weight copies whose
    total exceeds 2x L2, so every launch streams its weights from HBM.  Parity: sampled rows vs oracle.ref32."""
    cells = []
    nbytes = packed_bytes(fmt, o, k)
    if W is None:
        W = make_weights(torch, fmt, o, k, seed)
    copies = max(2, min(64, -(-2 * L2_BYTES // nbytes)))   # >= 2: consecutive launches never read the same bytes
    Ws = [W] + [W.clone() for _ in range(copies - 1)]
    pick = sample_rows(o, 96)
    for t in ts:
        X = torch.randn((t, k), device="cuda", dtype=torch.float16)
        C = torch.empty((t, o), device="cuda", dtype=torch.float16)
        f = ext.FMT_ID[fmt]
        C.zero_()
        ext.mm(f, Ws[-1], X, o, t, k, out=C)   # also the one-time setup (function attributes, workspaces) before capture
        torch.cuda.synchronize()
        par = parity_of(torch, C, pick, ref_rows(torch, fmt, W, k, pick, X.cpu().numpy()))
        n = max(16, copies * 2)
        g = graph_of(torch, [lambda i=i: ext.mm(f, Ws[i % copies], X, o, t, k, out=C) for i in range(n)])
        ms = timed(torch, None, g.replay, 5, 3, 1) / n
        gbs = nbytes / (ms * 1e-3) / 1e9
        cells.append({"cell": f"decode {name} T={t}", "family": "decode", "fmt": fmt, "O": o, "K": k, "T": t,
                      "us": round(ms * 1e3, 2), "achieved": round(gbs, 1), "unit": "GB/s", "peak": hbm_peak,
                      "frac": round(gbs / hbm_peak, 3), "frac_8TBps": round(gbs / 8000.0, 3),
                      "kernel": describe(ext, fmt, o, t, k), "weight_copies": copies, "parity": par})
        if gbs > hbm_peak:   # the peak is a COPY figure (read + write, with bus turnarounds): a read-only stream can pass it
            cells[-1]["peak_note"] = ("above the measured copy peak: MEASURED_PEAKS.json is read+write copy bandwidth; this "
                                      "kernel only reads (2 weight copies alternate, each larger than L2)")
        del g
    del Ws
    torch.cuda.empty_cache()
    return cells


def prefill_cell(torch, ext, name, fmt, o, k, t, tf_peak, W=None, seed=11):
    """Prefill TFLOP/s of one layer: eager launches (ms-scale kernels; W + X + C exceed L2).  Parity: sampled
    out-features (all T tokens) vs oracle.ref32."""
    if W is None:
        W = make_weights(torch, fmt, o, k, seed)
    X = torch.randn((t, k), device="cuda", dtype=torch.float16)
    C = torch.empty((t, o), device="cuda", dtype=torch.float16)
    f = ext.FMT_ID[fmt]
    ext.mm(f, W, X, o, t, k, out=C)
    torch.cuda.synchronize()
    pick = sample_rows(o, 48)
    par = parity_of(torch, C, pick, ref_rows(torch, fmt, W, k, pick, X.cpu().numpy()))
    ms = timed(torch, None, lambda: ext.mm(f, W, X, o, t, k, out=C), 6, 3, 1)
    tf = 2.0 * t * o * k / (ms * 1e-3) / 1e12
    cell = {"cell": f"prefill {name} T={t}", "family": "prefill", "fmt": fmt, "O": o, "K": k, "T": t,
            "us": round(ms * 1e3, 1), "achieved": round(tf, 1), "unit": "TFLOP/s", "peak": tf_peak,
            "frac": round(tf / tf_peak, 3), "frac_2250_nominal": round(tf / 2250.0, 3),
            "kernel": describe(ext, fmt, o, t, k), "parity": par}
    del W, X, C
    torch.cuda.empty_cache()
    return cell


def swiglu_cell(torch, ext, name, fmt, o, k, ts, hbm_peak, seed=17):
    """Fused SwiGLU up-projection (ggq_mm_swiglu): GB/s over the packed bytes of BOTH matrices, next to the two mmq calls
    + torch silu*mul it replaces.  Parity: sampled rows vs silu(fp16(ref32 gate)) * fp16(ref32 up)."""
    from kernels.swiglu import mmq_swiglu
    cells = []
    nbytes = 2 * packed_bytes(fmt, o, k)
    copies = max(1, min(32, -(-2 * L2_BYTES // nbytes)))
    Wg = [make_weights(torch, fmt, o, k, seed)]
    Wu = [make_weights(torch, fmt, o, k, seed + 1)]
    Wg += [Wg[0].clone() for _ in range(copies - 1)]
    Wu += [Wu[0].clone() for _ in range(copies - 1)]
    pick = sample_rows(o, 96)
    f = ext.FMT_ID[fmt]
    for t in ts:
        X = torch.randn((t, k), device="cuda", dtype=torch.float16)
        C = mmq_swiglu(fmt, Wg[0], Wu[0], X, o, t, k)
        torch.cuda.synchronize()
        x_np = X.cpu().numpy()
        g = ref_rows(torch, fmt, Wg[0], k, pick, x_np).astype(np.float16).astype(np.float32)
        u = ref_rows(torch, fmt, Wu[0], k, pick, x_np).astype(np.float16).astype(np.float32)
        par = parity_of(torch, C, pick, g / (1.0 + np.exp(-g)) * u)
        n = max(16, copies * 2)
        gr = graph_of(torch, [lambda i=i: mmq_swiglu(fmt, Wg[i % copies], Wu[i % copies], X, o, t, k) for i in range(n)])
        ms = timed(torch, None, gr.replay, 5, 3, 1) / n

        def unfused(i):
            a = ext.mm(f, Wg[i % copies], X, o, t, k)
            b = ext.mm(f, Wu[i % copies], X, o, t, k)
            return torch.nn.functional.silu(a) * b
        gr2 = graph_of(torch, [lambda i=i: unfused(i) for i in range(n)])
        ms2 = timed(torch, None, gr2.replay, 5, 3, 1) / n
        gbs = nbytes / (ms * 1e-3) / 1e9
        cells.append({"cell": f"swiglu {name} T={t}", "family": "decode (fused gate/up + silu*mul)", "fmt": fmt, "O": o, "K": k,
                      "T": t, "us": round(ms * 1e3, 2), "achieved": round(gbs, 1), "unit": "GB/s", "peak": hbm_peak,
                      "frac": round(gbs / hbm_peak, 3), "frac_8TBps": round(gbs / 8000.0, 3),
                      "us_two_mmq_plus_torch_silu_mul": round(ms2 * 1e3, 2), "weight_copies": copies, "parity": par})
        del gr, gr2
    del Wg, Wu
    torch.cuda.empty_cache()
    return cells


def q8_1_cell(torch, ext, name, fmt, o, k, t, hbm_peak, W):
    """The reference-arithmetic mode (Q8_1 activations, integer block dots, fp16 accumulator: ggq_mm_ref_q8_1), timed on
    the device; parity = BIT-IDENTICAL to the oracle's port of kernels/cpu_impls on sampled rows."""
    from kernels import q8_1_mode
    from oracle import ggq_oracle as orc
    from utils.quantize.q8_1 import quantize_to_q8_1
    fn = {"q8_0": q8_1_mode.mmq_q8_0_q8_1, "q4_k": q8_1_mode.mmq_q4_k_q8_1, "q6_k": q8_1_mode.mmq_q6_k_q8_1}[fmt]
    X = torch.randn((t, k), device="cuda", dtype=torch.float16)
    XQ = quantize_to_q8_1(X)
    W2 = W.clone()
    C = fn(W, XQ, o, t, k)
    torch.cuda.synchronize()
    pick = sample_rows(o, 48)
    rb = packed_bytes(fmt, 1, k)
    Wc = W.view(-1, rb)[torch.from_numpy(pick).to(W.device)].cpu().numpy().reshape(-1)
    want = np.ascontiguousarray(orc.mmq_cpu(fmt, Wc, X.cpu().numpy(), len(pick), t, k))
    got = np.ascontiguousarray(C[:, torch.from_numpy(pick).to(C.device)].cpu().numpy())
    same = bool(np.array_equal(got.view(np.uint16), want.view(np.uint16)))
    it = [0]

    def step():
        it[0] += 1
        fn(W if it[0] & 1 else W2, XQ, o, t, k)
    ms = timed(torch, None, step, 10, 3, 1)
    nbytes = packed_bytes(fmt, o, k)
    gbs = nbytes / (ms * 1e-3) / 1e9
    del W2
    torch.cuda.empty_cache()
    return {"cell": f"q8_1-mode {name} T={t}", "family": "reference arithmetic (Q8_1 activations, integer block dots: DP4A at T <= 2, tensor-core IMMA beyond; fp16 accumulator)",
            "fmt": fmt, "O": o, "K": k, "T": t, "us": round(ms * 1e3, 2), "achieved": round(gbs, 1), "unit": "GB/s",
            "peak": hbm_peak, "frac": round(gbs / hbm_peak, 3), "frac_8TBps": round(gbs / 8000.0, 3),
            "parity": {"bit_identical_to_cpu_impls_port": same, "rows_sampled": int(len(pick)), "ok": same}}


def single_gpu_cells(torch, ext, hbm_peak, tf_peak, W_head):
    cells = []
    # BASELINE configs[1] / [2] / [3]: decode, three quant types, M = 1..16
    cells += decode_cell(torch, ext, "Q4_K lm_head 128256x4096", "q4_k", 128256, 4096, (1, 8, 16), hbm_peak, W=W_head)
    cells += decode_cell(torch, ext, "Q6_K lm_head 128256x4096", "q6_k", 128256, 4096, (1, 8, 16), hbm_peak)
    cells += decode_cell(torch, ext, "Q8_0 FFN 28672x8192", "q8_0", 28672, 8192, (1, 8, 16), hbm_peak)
    # configs[0] (the reference's own CPU-runnable case) and the small Llama-3-8B layers
    cells += decode_cell(torch, ext, "Q8_0 4096x4096 [configs[0]]", "q8_0", 4096, 4096, (1,), hbm_peak)
    cells += decode_cell(torch, ext, "Q4_K 14336x4096", "q4_k", 14336, 4096, (1, 16), hbm_peak)
    cells += decode_cell(torch, ext, "Q4_K 4096x4096", "q4_k", 4096, 4096, (1,), hbm_peak)
    cells += decode_cell(torch, ext, "Q8_0 lm_head 128256x4096", "q8_0", 128256, 4096, (1,), hbm_peak)
    cells += decode_cell(torch, ext, "Q6_K down_proj 4096x14336", "q6_k", 4096, 14336, (1,), hbm_peak)
    # configs[3] / [2]: prefill through tcgen05
    cells.append(prefill_cell(torch, ext, "Q4_K FFN 28672x8192", "q4_k", 28672, 8192, 4096, tf_peak))
    cells.append(prefill_cell(torch, ext, "Q8_0 FFN 28672x8192", "q8_0", 28672, 8192, 4096, tf_peak))
    cells.append(prefill_cell(torch, ext, "Q6_K down_proj 4096x14336", "q6_k", 4096, 14336, 2048, tf_peak))
    cells.append(prefill_cell(torch, ext, "Q6_K lm_head 128256x4096", "q6_k", 128256, 4096, 2048, tf_peak))
    # the reference's own arithmetic (Q8_1 activations) on the headline layer
    cells.append(q8_1_cell(torch, ext, "Q4_K lm_head 128256x4096", "q4_k", 128256, 4096, 1, hbm_peak, W_head))
    cells.append(q8_1_cell(torch, ext, "Q4_K lm_head 128256x4096", "q4_k", 128256, 4096, 8, hbm_peak, W_head))
    # the fused SwiGLU up-projection on the Llama-3-8B / 70B FFN shapes
    cells += swiglu_cell(torch, ext, "Q4_K FFN gate+up 2x14336x4096", "q4_k", 14336, 4096, (1, 8), hbm_peak)
    cells += swiglu_cell(torch, ext, "Q4_K FFN gate+up 2x28672x8192", "q4_k", 28672, 8192, (1, 8), hbm_peak)
    # configs[4] layers on ONE GPU (the N = 1 point of the N-split cells below)
    cells += decode_cell(torch, ext, "cfg5 Q6_K lm_head 128256x8192", "q6_k", 128256, 8192, (1,), hbm_peak)
    cells += decode_cell(torch, ext, "cfg5 Q4_K FFN up 28672x8192", "q4_k", 28672, 8192, (1,), hbm_peak)
    cells += decode_cell(torch, ext, "cfg5 Q4_K FFN down 8192x28672", "q4_k", 8192, 28672, (1,), hbm_peak)
    return cells


# ---------------------------------------------------------------------------------------------
# N > 1: the N-split layer (multigpu/nsplit.py), exchanged output verified on every rank
# ---------------------------------------------------------------------------------------------
def gather_refs(torch, dist, fmt, W_shard, k, per, x_np, world, n=48):
    """Every rank computes the fp32 reference of sampled rows of ITS shard; all ranks receive all of them:
    (global column indices, ref[T, world * n])."""
    pick = sample_rows(per, n, seed=3)[:n]
    ref = ref_rows(torch, fmt, W_shard, k, pick, x_np)                       # [T, n]
    mine = torch.from_numpy(ref).to("cuda").contiguous()
    allr = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allr, mine)
    cols = np.concatenate([pick + r * per for r in range(world)])
    return cols, torch.cat(allr, dim=1).cpu().numpy()


def check_exchanged(torch, dist, C, cols, ref, what):
    """Tier-1 check of the exchanged [T, O] on THIS rank against the references of all shards; all ranks must pass."""
    par = parity_of(torch, C, cols, ref)
    ok = torch.tensor([1 if par["ok"] else 0], device="cuda", dtype=torch.int32)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    par["ok_all_ranks"] = bool(ok.item())
    par["columns_from_every_shard"] = True
    if not par["ok_all_ranks"]:
        raise AssertionError(f"{what}: exchanged output outside tolerance on some rank: {par}")
    return par


def nsplit_decode(torch, dist, nsplit, name, fmt, o, k, t, world, rank, exchange, steps, warmup, hbm_peak, W=None, seed=21):
    """Strong-scaled decode step of one N-split layer: graph-captured fused steps rotating through enough shard
    copies that consecutive steps cannot be served from L2.  The exchanged output of the eager steps AND of the
    graph replays is verified on every rank against references from every shard."""
    per = o // world
    if W is None:
        W = make_weights(torch, fmt, per, k, seed + rank)
    shard_bytes = packed_bytes(fmt, per, k)
    copies = 1 if shard_bytes > 1.2 * L2_BYTES else min(8, -(-2 * L2_BYTES // shard_bytes))
    x = torch.randn((t, k), dtype=torch.float16, generator=torch.Generator().manual_seed(5)).to("cuda")
    dist.broadcast(x, src=0)
    cols, ref = gather_refs(torch, dist, fmt, W, k, per, x.cpu().numpy(), world)
    layers = []
    for c in range(copies):
        layers.append(nsplit.NSplitLinear(fmt, W if c == 0 else W.clone(), o, k, mode=exchange, max_tokens=16))
    # the one-kernel fused decode step covers T <= 8 (and T*K <= 65536); more tokens take forward()'s GEMM path (tcgen05
    # skinny kernel with peer-stored tiles between two symmetric-memory barriers), issued eagerly
    fused = exchange == "fused" and layers[0].fused_decode_ok(t)
    pars = {}
    for L in layers:
        if fused and rank == 0:
            L.set_resident_input(x)
    # eager steps (also the warm-up of every copy), verified
    for i in range(3):
        for L in layers:
            out = L.forward(None, T=t) if fused else L.forward(x.clone())
    torch.cuda.synchronize()
    pars["eager"] = check_exchanged(torch, dist, out, cols, ref, name + " eager")
    if fused:
        gsteps = max(copies, 8 // copies * copies)
        torch.cuda.synchronize()
        dist.barrier()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            for i in range(gsteps):
                layers[i % copies]._launch_sync(t)

        def replay():
            g.replay()
            for i in range(gsteps):
                layers[i % copies]._epoch += 1
        reps = max(2, -(-steps // gsteps))
        ms = timed(torch, dist, replay, reps, max(1, -(-warmup // gsteps)), world) / gsteps
        out = layers[(gsteps - 1) % copies].last_output(t)
        pars["graph_replay"] = check_exchanged(torch, dist, out, cols, ref, name + " graph replay")
        launch = f"CUDA graph of {gsteps} fused steps over {copies} shard copies, replayed {reps}x"
        nsteps = reps * gsteps
    else:
        it = [0]

        def step():
            layers[it[0] % copies].forward(x)
            it[0] += 1
        ms = timed(torch, dist, step, steps, warmup, world)
        launch = ("eager: NCCL broadcast of X + skinny GEMM with peer-stored tiles + symmetric-memory barriers per step"
                  if exchange == "fused" else "eager NCCL broadcast + mm + all-gather per step")
        nsteps = steps
    nbytes = packed_bytes(fmt, o, k)
    gbs = nbytes / (ms * 1e-3) / 1e9
    cell = {"cell": f"nsplit x{world} decode {name} T={t}", "family": "decode+exchange", "fmt": fmt, "O": o, "K": k, "T": t,
            "us": round(ms * 1e3, 2), "achieved": round(gbs, 1), "unit": "GB/s (whole layer, all ranks)",
            "peak": hbm_peak * world, "frac": round(gbs / (hbm_peak * world), 3), "exchange": exchange, "launch": launch,
            "steps_timed": nsteps, "shard_copies": copies, "parity": pars}
    return cell, ms, layers, nsteps


def guarded(torch, name, fn):
    """An extra cell (not the headline): a parity failure — raised on every rank at the same point, check_exchanged
    all-reduces the verdict — is REPORTED in the cell list instead of taking the headline line down with it."""
    try:
        c = fn()
    except AssertionError as e:
        c = {"cell": name, "error": str(e)[:400], "parity": {"ok": False}}
    torch.cuda.empty_cache()
    return c


def nsplit_prefill(torch, dist, nsplit, name, fmt, o, k, t, world, rank, exchange, tf_peak, seed=31):
    per = o // world
    W = make_weights(torch, fmt, per, k, seed + rank)
    x = torch.randn((t, k), device="cuda", dtype=torch.float16)
    dist.broadcast(x, src=0)
    cols, ref = gather_refs(torch, dist, fmt, W, k, per, x.cpu().numpy(), world, n=16)
    L = nsplit.NSplitLinear(fmt, W, o, k, mode=exchange, max_tokens=t)
    xs = x.clone()
    out = L.forward(xs)
    torch.cuda.synchronize()
    par = check_exchanged(torch, dist, out, cols, ref, name + " prefill")
    ms = timed(torch, dist, lambda: L.forward(xs), 5, 3, world)
    tf = 2.0 * t * o * k / (ms * 1e-3) / 1e12
    cell = {"cell": f"nsplit x{world} prefill {name} T={t}", "family": "prefill+exchange", "fmt": fmt, "O": o, "K": k, "T": t,
            "us": round(ms * 1e3, 1), "achieved": round(tf, 1), "unit": "TFLOP/s (whole layer, all ranks)",
            "peak": tf_peak * world, "frac": round(tf / (tf_peak * world), 3), "exchange": exchange,
            "step": "NCCL broadcast of X + per-rank GEMM into the own buffer + one 2-D DMA copy per peer (copy engines, one side "
                    "stream each) between two symmetric-memory barriers" if exchange == "fused"
                    else "NCCL broadcast + GEMM + all-gather", "parity": par}
    del L, W
    torch.cuda.empty_cache()
    return cell


def run_ours(args):
    import torch
    import torch.distributed as dist

    from kernels import _ext as ext
    from kernels.mmq_q4_k import mmq_q4_k

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    ext.lib()
    hbm_peak, tf_peak, peak_src = peaks()
    f = ext.FMT_ID[FMT]

    rows = O // world
    assert rows * world == O
    W = make_weights(torch, FMT, rows, K, 1234 + rank)
    total_bytes = packed_bytes(FMT, O, K)
    k_bytes = packed_bytes(FMT, rows, K)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    out = {}

    if world == 1:
        # ---------------- headline, one GPU ----------------
        W2 = W.clone()   # steps alternate between two copies: 591 MB streamed between two uses of the same bytes
        Wc = (W, W2)
        x_host = [torch.randn((T, K), dtype=torch.float16).pin_memory() for _ in range(2)]
        c_host = [torch.empty((T, O), dtype=torch.float16).pin_memory() for _ in range(2)]
        x_dev = [h.to("cuda") for h in x_host]
        c_dev = [torch.empty((T, O), device="cuda", dtype=torch.float16)]
        pick = sample_rows(rows, 128)
        ref = ref_rows(torch, FMT, W, K, pick, x_host[0].numpy())
        C = mmq_q4_k(W, x_dev[0], rows, T, K)
        C = mmq_q4_k(W2, x_dev[0], rows, T, K)
        torch.cuda.synchronize()
        parity = parity_of(torch, C, pick, ref)
        assert parity["ok"], ("bench parity", parity)

        it = [0]

        def step_device():
            i = it[0] = it[0] + 1
            ext.mm(f, Wc[i & 1], x_dev[0], rows, T, K, out=c_dev[0])

        l0 = ext.launch_count()
        ms = timed(torch, dist, step_device, args.steps, args.warmup, world)
        launches = (ext.launch_count() - l0) - args.warmup
        ms_k = ms   # the step IS the kernel (one library call per step)

        # e2e: the host-buffer call (kernels.host.HostPipe -> ggq_mm_host): pinned host X in, host C out, every step; the
        # library runs copy-in / kernel / copy-out on three streams with rotating slots, so the copies of neighbouring
        # steps overlap the kernel and the kernels stay back to back
        from kernels.host import HostPipe
        from kernels import host as _host
        pipe = HostPipe(FMT, Wc[0], rows, K, max_tokens=T, depth=3)   # (the weights are an argument of every call)
        mm_host = _host._lib().ggq_mm_host

        def step_e2e():
            i = it[0] = it[0] + 1
            b = i & 1
            rc = mm_host(pipe._h, f, Wc[b].data_ptr(), x_host[b].data_ptr(), c_host[b].data_ptr(), rows, T, K)
            assert rc == 0, rc

        def timed_e2e(steps, warmup):
            for _ in range(warmup):
                step_e2e()
            pipe.sync()
            torch.cuda.synchronize()
            s_in, s_out = pipe.stream(0), pipe.stream(2)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(s_in)            # the first copy-in of the region follows on this stream
            for _ in range(steps):
                step_e2e()
            e1.record(s_out)           # after the last copy-out
            pipe.sync()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / steps
        ms_e2e = timed_e2e(args.steps, args.warmup)
        e2e_par = parity_of(torch, c_host[it[0] & 1].to("cuda"), pick,
                            ref_rows(torch, FMT, W, K, pick, x_host[it[0] & 1].numpy()))
        assert e2e_par["ok"], ("e2e parity", e2e_par)
        # latency form: the public entry point, one step at a time, host synchronisation after every step
        def step_sync():
            x_dev[0].copy_(x_host[0], non_blocking=True)
            c_host[0].copy_(mmq_q4_k(W, x_dev[0], rows, T, K), non_blocking=True)
            torch.cuda.synchronize()
        for _ in range(3):
            step_sync()
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step_sync()
        ms_lat = (time.perf_counter() - t0) / args.steps * 1e3

        # isolated launches: a device synchronisation + a foreign kernel in front, no overlap with a predecessor
        iso = []
        for _ in range(12):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda._sleep(400_000)
            e0.record()
            ext.mm(f, Wc[len(iso) & 1], x_dev[0], rows, T, K, out=c_dev[0])
            e1.record()
            torch.cuda.synchronize()
            iso.append(e0.elapsed_time(e1))
        ms_iso = sorted(iso)[len(iso) // 2]
        # ~0.5 s of the same step back to back, so the clock sampler sees the GPU under this load more than once
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < 0.5:
            for _ in range(200):
                step_device()
            torch.cuda.synchronize()
        clocks = sampler.stop() if sampler else None
        cfg = {"workload": WORKLOAD, "l2": "steps alternate between two copies of the packed weights (2 x 295.5 MB vs 126 MB L2): "
               "every step streams its weights from HBM", "parallelism": "single GPU", "launch": "one library call per step",
               "weights": "fp32 N(0, 0.02) packed by utils/quantize (the repo's GPU packers, byte-identical to the reference's)",
               "fmt": FMT, "O": O, "K": K, "T": T}
        e2e = {"value": total_bytes / (ms_e2e * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": ms_e2e,
               "h2d_bytes_per_step": T * K * 2, "d2h_bytes_per_step": T * O * 2,
               "mode": "throughput: ggq_mm_host (kernels.host.HostPipe) per step — pinned H2D of X, kernel, D2H of C on the "
                       "library's three streams with rotating slots; timed with CUDA events from the first copy-in to the last copy-out",
               "latency_ms_per_step_synchronised": ms_lat, "parity": e2e_par}
        extra_cells = []
    else:
        # ---------------- headline, N-split ----------------
        from multigpu import nsplit
        exchange = args.exchange
        try:
            probe = nsplit.NSplitLinear(FMT, W, O, K, mode=exchange, max_tokens=16)
            del probe
        except Exception as e:  # no peer-mappable (symmetric) memory on this box: NCCL exchange instead
            if exchange != "fused":
                raise
            print(f"[bench] fused exchange unavailable ({type(e).__name__}: {e}); using NCCL", file=sys.stderr)
            exchange = "nccl"
        l0 = ext.launch_count()
        cell, ms, layers, nsteps = nsplit_decode(torch, dist, nsplit, "Q4_K lm_head 128256x4096", FMT, O, K, T, world, rank,
                                                 exchange, args.steps, args.warmup, hbm_peak, W=W)
        launches = nsteps
        parity = cell["parity"]
        # e2e: rank 0 copies X from pinned host memory into the peer-visible buffer, every rank copies the
        # exchanged [T, O] back to its host
        x_host = torch.randn((T, K), dtype=torch.float16).pin_memory()
        c_host = torch.empty((T, O), dtype=torch.float16).pin_memory()
        x_dev = x_host.to("cuda")
        L0 = layers[0]

        def step_e2e():
            if exchange == "fused":
                if rank == 0:
                    L0.input_buffer(T).copy_(x_host, non_blocking=True)
                c_host.copy_(L0.forward(None, T=T), non_blocking=True)
            else:
                x_dev.copy_(x_host, non_blocking=True)
                c_host.copy_(L0.forward(x_dev), non_blocking=True)
        ms_e2e = timed(torch, dist, step_e2e, args.steps, args.warmup, world)
        # the shard kernel alone (no exchange), for the roofline of the dominant kernel
        c_shard = torch.empty((T, rows), device="cuda", dtype=torch.float16)
        Wk = [L.A for L in layers]
        it = [0]

        def step_kernel():
            i = it[0] = it[0] + 1
            ext.mm(f, Wk[i % len(Wk)], x_dev, rows, T, K, out=c_shard)
        ms_k = timed(torch, dist, step_kernel, args.steps, args.warmup, world)
        ms_iso = None
        t0 = time.perf_counter()
        while time.perf_counter() - t0 < 0.5:
            for _ in range(200):
                step_kernel()
            torch.cuda.synchronize()
        clocks = sampler.stop() if sampler else None
        cfg = {"workload": WORKLOAD, "l2": f"steps rotate through {cell['shard_copies']} copies of the rank's "
               f"{k_bytes / 1e6:.1f} MB shard (L2 is 126 MB): every step streams its weights from HBM",
               "parallelism": f"N-split x{world}, exchange={exchange}", "launch": cell["launch"],
               "weights": "fp32 N(0, 0.02) packed by utils/quantize (the repo's GPU packers)",
               "fmt": FMT, "O": O, "K": K, "T": T}
        e2e = {"value": total_bytes / (ms_e2e * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": ms_e2e,
               "h2d_bytes_per_step": T * K * 2, "d2h_bytes_per_step": T * O * 2 * world,
               "mode": "eager Python call per step on every rank: rank 0 H2D of X into the peer-visible buffer, fused step, "
                       "D2H of the exchanged [T, O] on every rank"}
        extra_cells = []
        del layers, L0
        torch.cuda.empty_cache()
        if not args.no_cells:
            # BASELINE configs[4]: Llama-3-70B-class Q4_K FFN + Q6_K lm_head, column-sharded
            for (nm, fm, oo, kk) in (("cfg5 Q6_K lm_head 128256x8192", "q6_k", 128256, 8192),
                                     ("cfg5 Q4_K FFN up 28672x8192", "q4_k", 28672, 8192),
                                     ("cfg5 Q4_K FFN down 8192x28672", "q4_k", 8192, 28672)):
                for tt in (1, 16):
                    def one(nm=nm, fm=fm, oo=oo, kk=kk, tt=tt):
                        c, _, ls, _ = nsplit_decode(torch, dist, nsplit, nm, fm, oo, kk, tt, world, rank, exchange, 40, 8, hbm_peak)
                        del ls
                        return c
                    extra_cells.append(guarded(torch, f"nsplit x{world} decode {nm} T={tt}", one))
            extra_cells.append(guarded(torch, f"nsplit x{world} prefill cfg5 Q4_K FFN up T=4096", lambda: nsplit_prefill(
                torch, dist, nsplit, "cfg5 Q4_K FFN up 28672x8192", "q4_k", 28672, 8192, 4096, world, rank, exchange, tf_peak)))
            extra_cells.append(guarded(torch, f"nsplit x{world} prefill cfg5 Q6_K lm_head T=2048", lambda: nsplit_prefill(
                torch, dist, nsplit, "cfg5 Q6_K lm_head 128256x8192", "q6_k", 128256, 8192, 2048, world, rank, exchange, tf_peak)))

    cells = []
    if rank == 0 and world == 1 and not args.no_cells:
        cells = single_gpu_cells(torch, ext, hbm_peak, tf_peak, W)
    cells += extra_cells

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    value = total_bytes / (ms * 1e-3) / 1e9
    achieved = k_bytes / (ms_k * 1e-3) / 1e9
    kernel = describe(ext, FMT, rows, T, K)
    side = traffic_sidecar().get(f"{FMT} O={rows} K={K} T={T}") if world == 1 else None
    cpu_threads = 1
    cpu_gbs, cpu_dt, cpu_rows = cpu_path_gbs(65536, 1, 0, cpu_threads)
    roof = {"bound": "hbm", "kernel": kernel, "achieved": achieved, "peak": hbm_peak, "peak_source": peak_src, "unit": "GB/s",
            "frac": achieved / hbm_peak, "frac_of_8TBps_nominal": achieved / 8000.0, "us_per_launch": ms_k * 1e3,
            "algorithmic_bytes_per_launch": k_bytes,
            "timing": "CUDA events around the K back-to-back launches of the timed region (launches overlap their prologue "
                      "with the previous kernel's drain); `isolated` = median of single launches separated by a device "
                      "synchronisation",
            "traffic": None if not side else side["dram_read_bytes"] + side["dram_write_bytes"],
            "traffic_source": None if not side else side["source"]}
    if ms_iso is not None:
        roof["isolated"] = {"us_per_launch": ms_iso * 1e3, "achieved": k_bytes / (ms_iso * 1e-3) / 1e9,
                            "frac": k_bytes / (ms_iso * 1e-3) / 1e9 / hbm_peak}
    out = {
        "metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f16", "arithmetic": "fp16 activations x in-register dequantized weights, fp32 accumulate, fp16 out",
        "data": "synthetic", "config": cfg, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roof,
        "cpu_baseline": {"value": cpu_gbs, "unit": "GB/s", "cores": cpu_threads, "kind": "port",
                         "sample": f"{cpu_rows} of {O} rows, one pass ({cpu_dt:.1f} s); {PORT_NOTE}"},
        "parity": parity, "cells": cells,
    }
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cells", action="store_true", help="headline only (skip the per-cell table)")
    ap.add_argument("--detail", action="store_true", help="(kept for compatibility: the cells are always reported)")
    ap.add_argument("--exchange", default="fused", choices=["nccl", "fused"],
                    help="N>1: how the [T, O/N] slices reach every rank (NCCL all-gather, or peer stores fused into the kernel)")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
