mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r49_pytest.log; cat gpurun_out/r49_pytest.log
timeout 300 python tools/jac_probe.py 500x150 630x300 4096x1024 > gpurun_out/r49_jac_probe.log 2>&1; cut -c1-330 gpurun_out/r49_jac_probe.log
