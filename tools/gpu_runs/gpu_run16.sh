set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r16_pytest.log; cat gpurun_out/r16_pytest.log
timeout 600 python tools/jac_probe.py 2048x512 4096x1024 > gpurun_out/r16_jac_probe.log 2>&1; cat gpurun_out/r16_jac_probe.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_jac_assemble_rt" -s 1 -c 1 -o gpurun_out/r16_assemble_full python tools/jac_probe.py 2048x512 > gpurun_out/r16_ncu_jac.log 2>&1; tail -n 3 gpurun_out/r16_ncu_jac.log
ls -la gpurun_out
