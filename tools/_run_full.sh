set -x
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -c 2500 gpurun_out/bench_default.json
timeout 900 python bench.py --detail > gpurun_out/bench_detail.json 2> gpurun_out/bench_detail.err; echo detail rc=$?
