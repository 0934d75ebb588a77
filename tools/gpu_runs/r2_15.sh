mkdir -p gpurun_out
timeout 120 python tools/res_one.py 96x48 6 1 2>&1 | tail -2 | tee gpurun_out/r2_15_first.log
if grep -q "variant 6 ms" gpurun_out/r2_15_first.log; then
  timeout 300 compute-sanitizer --tool memcheck --print-limit 3 python tools/res_one.py 70x21 6 1 2>&1 | grep -v "^=========     \(Host\|    \)" | head -20 | tee gpurun_out/r2_15_memcheck.log
  for v in 4 6; do timeout 120 python tools/res_one.py 8192x2048 $v 10; done 2>&1 | grep variant | tee gpurun_out/r2_15_times.log
  for s in 1024x2048 630x300 500x150; do for v in 4 6; do timeout 120 python tools/res_one.py $s $v 20; done; done 2>&1 | grep variant | tee -a gpurun_out/r2_15_times.log
  timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_residual_fast_bulk -c 1 -o gpurun_out/r2_15_bulk python tools/res_one.py 8192x2048 6 2 > gpurun_out/r2_15_ncu.log 2>&1
  timeout 1500 python -m pytest tests -m gpu -q -x --deselect tests/test_multigpu_gpu.py 2>&1 | tail -8 | tee gpurun_out/r2_15_pytest.log
  timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_15_bench.json 2> gpurun_out/r2_15_bench.err; tail -c 3500 gpurun_out/r2_15_bench.json; tail -3 gpurun_out/r2_15_bench.err
fi
