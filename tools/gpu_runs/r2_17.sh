mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_scheme_orders_gpu.py -q > gpurun_out/r2_17_pytest.log 2>&1
tail -30 gpurun_out/r2_17_pytest.log
