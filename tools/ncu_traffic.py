#!/usr/bin/env python
"""profiles/ncu_traffic.json from the committed `ncu --set full` summaries (profiles/*_ncu_summary.csv, written by
tools/ncu_summary.py): per-launch DRAM bytes of the kernels bench.py reports, keyed "fmt O=.. K=.. T=..".
bench.py reads the file for `roofline.traffic` (null when a cell has no capture)."""
import csv
import json
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
# summary file -> workload key
CAPTURES = {
    "r2_decode_q4k_T1_lmhead_ncu_summary.csv": "q4_k O=128256 K=4096 T=1",
    "r2_decode_q6k_T8_lmhead_ncu_summary.csv": "q6_k O=128256 K=4096 T=8",
    "r2_decode_q8_0_T1_ffn_ncu_summary.csv": "q8_0 O=28672 K=8192 T=1",
    "r2_skinny_q8_0_T16_ffn_ncu_summary.csv": "q8_0 O=28672 K=8192 T=16",
    "r2_skinny_q4k_T16_lmhead_ncu_summary.csv": "q4_k O=128256 K=4096 T=16",
    "r2_prefill_q4k_ffn_ncu_summary.csv": "q4_k O=28672 K=8192 T=4096",
    "r2_prefill_q8_0_ffn_ncu_summary.csv": "q8_0 O=28672 K=8192 T=4096",
    "r2_prefill_q6k_lmhead_ncu_summary.csv": "q6_k O=128256 K=4096 T=2048",
}


def main():
    out = {}
    for name, key in CAPTURES.items():
        path = os.path.join(ROOT, "profiles", name)
        if not os.path.exists(path):
            continue
        m = {r["metric"]: (r["unit"], r["value"]) for r in csv.DictReader(open(path))}

        def b(metric):
            u, v = m[metric]
            return int(float(v) * UNIT[u])
        us_u, us_v = m["gpu__time_duration.sum"]
        out[key] = {"dram_read_bytes": b("dram__bytes_read.sum"), "dram_write_bytes": b("dram__bytes_write.sum"),
                    "ncu_us": float(us_v) * {"us": 1.0, "ms": 1e3, "ns": 1e-3}[us_u], "kernel": m["kernel"][1],
                    "source": f"profiles/{name} (ncu --set full, one launch, cold cache)"}
    with open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
    print(json.dumps(out, indent=1, sort_keys=True))


if __name__ == "__main__":
    main()
