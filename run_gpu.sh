timeout 1200 python -m pytest tests/test_long_run_statistics.py tests/test_gpu_parity.py -m gpu -q -s -k "ten_thousand or vv_split" 2>&1 | tail -30 > gpurun_out/pytest_gpu_8.log
