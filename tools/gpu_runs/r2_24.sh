mkdir -p gpurun_out
for v in 4 6; do timeout 120 python tools/res_one.py 8192x2048 $v 10; done 2>&1 | grep variant | tee gpurun_out/r2_24_times.log
for s in 1024x2048; do for v in 4 6; do timeout 120 python tools/res_one.py $s $v 20; done; done 2>&1 | grep variant | tee -a gpurun_out/r2_24_times.log
timeout 600 python -m pytest tests/test_residual_bulk_gpu.py tests/test_configs_gpu.py -q -x -k "not jacobian and not dz" 2>&1 | tail -4 | tee gpurun_out/r2_24_pytest.log
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_residual_fast_bulk -c 1 -o gpurun_out/r2_24_bulk python tools/res_one.py 8192x2048 6 2 > gpurun_out/r2_24_ncu.log 2>&1
