set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r25_pytest.log; cat gpurun_out/r25_pytest.log
timeout 300 python tools/jac_probe.py 500x150 630x300 2048x512 4096x1024 > gpurun_out/r25_jac_probe.log 2>&1; cat gpurun_out/r25_jac_probe.log
BROADCAST_B200_NO_GRAPH=1 timeout 300 python tools/jac_probe.py 500x150 4096x1024 > gpurun_out/r25_jac_probe_nograph.log 2>&1; cat gpurun_out/r25_jac_probe_nograph.log
