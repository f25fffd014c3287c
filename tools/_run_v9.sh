timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -k "chain or sync" 2>&1 | tail -5
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; python -c "
import json; d=json.loads(open('gpurun_out/bench_default.json').read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','e2e','gpu_launches')}); print(d['roofline'])"
