set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "residual or slab or streamed" 2>&1 | tail -25 > gpurun_out/r20_pytest.log; cat gpurun_out/r20_pytest.log
timeout 300 python tools/res_probe.py 2048x512 8192x2048 > gpurun_out/r20_res_probe.log 2>&1; cat gpurun_out/r20_res_probe.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_residual_fast -s 3 -c 1 -o gpurun_out/r20_residual_full python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-jacobian > gpurun_out/r20_ncu_res.log 2>&1; tail -n 3 gpurun_out/r20_ncu_res.log
