mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "dz_tangent_colour_loop or dz_colour_loop" 2>&1 | tail -12 | tee gpurun_out/r69_pytest.log
