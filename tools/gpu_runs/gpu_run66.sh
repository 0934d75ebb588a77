mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "dz_tangent" 2>&1 | tail -3 | tee gpurun_out/r66_pytest.log
timeout 200 python tools/dz_tangent_probe.py > gpurun_out/r66_dz_tangent_pairs.jsonl 2>/dev/null; cat gpurun_out/r66_dz_tangent_pairs.jsonl
timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_dz_tangent -s 3 -c 1 -f -o gpurun_out/r66_dz_tangent_pairs_full python tools/dz_tangent_probe.py 4096 1024 2 > gpurun_out/r66_ncu.log 2>&1; tail -n 1 gpurun_out/r66_ncu.log
