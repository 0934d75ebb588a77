set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r8_pytest.log; cat gpurun_out/r8_pytest.log
timeout 900 python tools/jac_probe.py 500x150 2048x512 4096x1024 > gpurun_out/r8_jac_probe.log 2>&1; cat gpurun_out/r8_jac_probe.log
timeout 900 python bench.py > gpurun_out/r8_bench.json 2>gpurun_out/r8_bench.err; cat gpurun_out/r8_bench.json; tail -5 gpurun_out/r8_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r8_jac_launches.csv python tools/jac_probe.py 4096x1024 > gpurun_out/r8_jac_launches.log 2>&1; tail -2 gpurun_out/r8_jac_launches.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_residual_tile -s 3 -c 1 -o gpurun_out/r8_residual_full python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-jacobian > gpurun_out/r8_ncu_res.log 2>&1; tail -3 gpurun_out/r8_ncu_res.log
ls -la gpurun_out
