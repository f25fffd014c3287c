timeout 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
timeout 200 python tools/probe_variant.py 2>&1
