mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_bc_general_gpu.py -q 2>&1 | tail -5
