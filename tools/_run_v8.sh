timeout 300 python -m pytest tests/test_gpu_parity.py -x -q 2>&1 | tail -3
echo "== PDL"; TS=1,8 timeout 300 python tools/probe_graph.py 2>&1
echo "== no PDL"; GGQ_NO_PDL=1 TS=1 timeout 300 python tools/probe_graph.py 2>&1
timeout 200 python tools/probe_variant.py 2>&1 | grep "128256\|28672"
