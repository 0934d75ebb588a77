mkdir -p gpurun_out
echo "concurrent: $(timeout 200 python tools/strips_only.py 8192x2048 2>&1 | tail -1)" | tee gpurun_out/r2_41_strips.log
echo "serial: $(BROADCAST_B200_STRIPS_SERIAL=1 timeout 200 python tools/strips_only.py 8192x2048 2>&1 | tail -1)" | tee -a gpurun_out/r2_41_strips.log
echo "concurrent 1024x2048: $(timeout 200 python tools/strips_only.py 1024x2048 2>&1 | tail -1)" | tee -a gpurun_out/r2_41_strips.log
echo "serial 1024x2048: $(BROADCAST_B200_STRIPS_SERIAL=1 timeout 200 python tools/strips_only.py 1024x2048 2>&1 | tail -1)" | tee -a gpurun_out/r2_41_strips.log
timeout 900 python -m pytest tests/test_strips_window_gpu.py tests/test_parity_gpu.py tests/test_configs_gpu.py tests/test_two_zone_gpu.py -q -m gpu -k "strip or window or hybrid or jacobian or banded or csr or two_zone" 2>&1 | tail -4
