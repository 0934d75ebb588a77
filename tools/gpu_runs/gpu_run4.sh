set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r4_pytest.log; cat gpurun_out/r4_pytest.log
timeout 900 python tools/jac_probe.py 500x150 630x300 2048x512 4096x1024 > gpurun_out/r4_jac_probe.log 2>&1; cat gpurun_out/r4_jac_probe.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r4_jac_launches.csv python tools/jac_probe.py 2048x512 > gpurun_out/r4_jac_launches.log 2>&1; tail -2 gpurun_out/r4_jac_launches.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_jac_assemble_rt|k_face_packages" -s 2 -c 2 -o gpurun_out/r4_facejac_full python tools/jac_probe.py 2048x512 > gpurun_out/r4_ncu_jac.log 2>&1; tail -3 gpurun_out/r4_ncu_jac.log
ls -la gpurun_out
