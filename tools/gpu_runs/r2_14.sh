mkdir -p gpurun_out
timeout 120 python tools/res_one.py 96x48 6 1 2>&1 | tail -2 | tee gpurun_out/r2_14_first.log
if grep -q "variant 6 ms" gpurun_out/r2_14_first.log; then
  timeout 300 compute-sanitizer --tool memcheck --print-limit 3 python tools/res_one.py 70x21 6 1 2>&1 | grep -v "^=========     \(Host\|    \)" | head -20 | tee gpurun_out/r2_14_memcheck.log
  timeout 900 python -m pytest tests/test_residual_bulk_gpu.py tests/test_configs_gpu.py -q -x 2>&1 | tail -5 | tee gpurun_out/r2_14_pytest.log
  for v in 4 6; do timeout 120 python tools/res_one.py 8192x2048 $v 10; done 2>&1 | grep variant | tee gpurun_out/r2_14_times.log
  for s in 1024x2048 630x300 500x150; do for v in 4 6; do timeout 120 python tools/res_one.py $s $v 20; done; done 2>&1 | grep variant | tee -a gpurun_out/r2_14_times.log
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_residual_fast_bulk -c 1 -o gpurun_out/r2_14_bulk python tools/res_one.py 8192x2048 6 2 > gpurun_out/r2_14_ncu.log 2>&1
  timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_14_bench.json 2> gpurun_out/r2_14_bench.err; tail -c 3000 gpurun_out/r2_14_bench.json; tail -3 gpurun_out/r2_14_bench.err
fi
