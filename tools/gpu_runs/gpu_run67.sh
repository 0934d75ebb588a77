mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "isothermal_wall_and_symmetry" 2>&1 | tail -12 | tee gpurun_out/r67_pytest_bcs.log
