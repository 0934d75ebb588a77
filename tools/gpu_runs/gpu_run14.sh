set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r14_pytest.log; cat gpurun_out/r14_pytest.log
timeout 300 python tools/res_probe.py 500x150 2048x512 8192x2048 > gpurun_out/r14_res_probe.log 2>&1; cat gpurun_out/r14_res_probe.log
BROADCAST_B200_RESIDUAL_NO_TMA=1 timeout 300 python tools/res_probe.py 8192x2048 > gpurun_out/r14_res_probe_notma.log 2>&1; cat gpurun_out/r14_res_probe_notma.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_residual_fast -s 3 -c 1 -o gpurun_out/r14_residual_full python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-jacobian > gpurun_out/r14_ncu_res.log 2>&1; tail -n 3 gpurun_out/r14_ncu_res.log
ls -la gpurun_out
