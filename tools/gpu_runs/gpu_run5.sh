set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r5_pytest.log; cat gpurun_out/r5_pytest.log
timeout 900 python tools/jac_probe.py 500x150 2048x512 4096x1024 > gpurun_out/r5_jac_probe.log 2>&1; cat gpurun_out/r5_jac_probe.log
timeout 600 python bench.py --no-e2e --no-cpu-baseline > gpurun_out/r5_bench.json 2>gpurun_out/r5_bench.err; cat gpurun_out/r5_bench.json; tail -3 gpurun_out/r5_bench.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_jac_assemble_rt" -s 1 -c 1 -o gpurun_out/r5_assemble_full python tools/jac_probe.py 2048x512 > gpurun_out/r5_ncu_jac.log 2>&1; tail -3 gpurun_out/r5_ncu_jac.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_residual_tile -s 3 -c 1 -o gpurun_out/r5_residual_full python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r5_ncu_res.log 2>&1; tail -3 gpurun_out/r5_ncu_res.log
ls -la gpurun_out
