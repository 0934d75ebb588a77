set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r13_pytest.log; cat gpurun_out/r13_pytest.log
timeout 600 python tools/res_probe.py 2048x512 8192x2048 > gpurun_out/r13_res_probe.log 2>&1; cat gpurun_out/r13_res_probe.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_residual_fast -s 3 -c 1 -o gpurun_out/r13_residual_full python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-jacobian > gpurun_out/r13_ncu_res.log 2>&1; tail -n 3 gpurun_out/r13_ncu_res.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r13_jac_launches.csv python tools/jac_probe.py 4096x1024 > gpurun_out/r13_jac_launches.log 2>&1; tail -n 2 gpurun_out/r13_jac_launches.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_jac_assemble_rt" -s 1 -c 1 -o gpurun_out/r13_assemble_full python tools/jac_probe.py 2048x512 > gpurun_out/r13_ncu_jac.log 2>&1; tail -n 3 gpurun_out/r13_ncu_jac.log
ls -la gpurun_out
