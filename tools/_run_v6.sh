timeout 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -5
timeout 200 python tools/probe_variant.py 2>&1
TS=1,4,8 timeout 300 python tools/probe_graph.py 2>&1
