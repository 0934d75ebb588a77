mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r58_pytest.log
timeout 200 python tools/dz_tangent_probe.py > gpurun_out/r58_dz_tangent.jsonl 2> gpurun_out/r58_dz_tangent.err; cat gpurun_out/r58_dz_tangent.jsonl; tail -n 3 gpurun_out/r58_dz_tangent.err
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_dz_tangent -s 3 -c 1 -f -o gpurun_out/r58_dz_tangent_full python tools/dz_tangent_probe.py 4096 1024 2 > gpurun_out/r58_ncu.log 2>&1; tail -n 2 gpurun_out/r58_ncu.log
