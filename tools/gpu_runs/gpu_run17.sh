set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "jacobian or face_lin or csr" 2>&1 | tail -25 > gpurun_out/r17_pytest.log; cat gpurun_out/r17_pytest.log
timeout 600 python tools/jac_probe.py 2048x512 4096x1024 > gpurun_out/r17_jac_probe.log 2>&1; cat gpurun_out/r17_jac_probe.log
