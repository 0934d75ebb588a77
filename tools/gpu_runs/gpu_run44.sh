mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 300 python tools/jac_probe.py 500x150 630x300 2>&1 | cut -c1-330
BROADCAST_B200_COO_GENERIC=1 timeout 300 python tools/jac_probe.py 500x150 2>&1 | cut -c1-330
