"""Prefill TFLOP/s per quant type on the BASELINE config 3/4 shapes (dev probe; GGQ_PREFILL_1CTA=1 selects the 1-CTA kernel)."""
import os, sys, json, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "gguf-triton-kernel_b200"))
import bench
from kernels import _ext as ext
_, tf_peak, _ = bench.peaks()
for r in bench.prefill_table(torch, ext, tf_peak):
    print(os.environ.get("GGQ_PREFILL_1CTA", "0"), json.dumps(r))
