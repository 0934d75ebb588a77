mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r50_pytest.log; cat gpurun_out/r50_pytest.log
timeout 300 python tools/jac_probe.py 500x150 630x300 4096x1024 > gpurun_out/r50_jac_probe.log 2>&1; cut -c1-330 gpurun_out/r50_jac_probe.log
BROADCAST_B200_STRIP_CHAINS=1 timeout 300 python tools/jac_probe.py 500x150 2>&1 | cut -c1-200 | sed "s/^/chains=1 /"
