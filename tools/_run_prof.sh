python bench.py --steps 20 --warmup 5 > gpurun_out/b20.json 2> gpurun_out/b20.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1b_bench_launches.csv python bench.py --steps 20 --warmup 5 > gpurun_out/ncu_launch.log 2>&1
python tools/ncu_one.py && ncu --set full --clock-control none --import-source on -k regex:decode_kernel --launch-skip 3 -c 1 -o gpurun_out/r1b_q4k_t1 -f python tools/ncu_one.py > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
