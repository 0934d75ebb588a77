mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12 > gpurun_out/r47_pytest.log; cat gpurun_out/r47_pytest.log
timeout 300 python tools/jac_probe.py 500x150 630x300 4096x1024 > gpurun_out/r47_jac_probe.log 2>&1; cut -c1-330 gpurun_out/r47_jac_probe.log
BROADCAST_B200_TANGENT_TILE=0 timeout 300 python tools/jac_probe.py 500x150 4096x1024 2>&1 | cut -c1-330 | sed "s/^/tile=0 /"
