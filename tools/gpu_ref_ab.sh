python -m pytest tests/test_gpu_quantize_ops.py -q -m gpu -k reference_arithmetic > gpurun_out/r2_ref4_pytest.log 2>&1; echo rc=$?; tail -3 gpurun_out/r2_ref4_pytest.log
python tools/dev_ref_time.py 1 2 4 8 16 > gpurun_out/r2_ref4_time.log 2>&1
GGQ_REFMODE_MMA_MIN_T=99 python tools/dev_ref_time.py 2 4 8 >> gpurun_out/r2_ref4_time.log 2>&1
cat gpurun_out/r2_ref4_time.log
