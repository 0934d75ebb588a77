mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_solve_gpu.py -q > gpurun_out/r2_23_pytest.log 2>&1
tail -40 gpurun_out/r2_23_pytest.log | cut -c1-300
