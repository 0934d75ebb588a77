mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r37_pytest.log; cat gpurun_out/r37_pytest.log
timeout 600 python tools/csr_probe.py 500x150 630x300 2048x512 > gpurun_out/r37_csr_probe.log 2>&1; cat gpurun_out/r37_csr_probe.log
