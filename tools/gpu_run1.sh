#!/bin/bash
# round-2 GPU pass: tests, the bench line, full ncu captures of the T=16 skinny kernels
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest.log 2>&1; echo "pytest rc=$?"
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench1.json 2> gpurun_out/r2_bench1.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2_bench1.err
NCU="ncu --set full --clock-control none --import-source on --launch-skip 4 --launch-count 1"
FMT=q4_k T=16 timeout 300 $NCU -k regex:skinny -o gpurun_out/r2_skinny_q4k_t16 -f python tools/ncu_one.py > gpurun_out/r2_ncu1.log 2>&1; echo "ncu1 rc=$?"
FMT=q6_k T=16 timeout 300 $NCU -k regex:skinny -o gpurun_out/r2_skinny_q6k_t16 -f python tools/ncu_one.py > gpurun_out/r2_ncu2.log 2>&1; echo "ncu2 rc=$?"
