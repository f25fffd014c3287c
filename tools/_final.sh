set -x
timeout 600 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; echo bench rc=$?
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo ref rc=$?; tail -c 600 gpurun_out/bench_ref.json
python bench.py --steps 20 --warmup 5 > gpurun_out/b20.json 2> gpurun_out/b20.err && ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1c_bench_launches.csv python bench.py --steps 20 --warmup 5 > gpurun_out/ncu_launch.log 2>&1; echo ncu rc=$?
