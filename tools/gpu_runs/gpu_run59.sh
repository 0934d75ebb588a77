mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_parity_gpu.py -m gpu -x -q -k "dz_tangent" 2>&1 | tail -4 | tee gpurun_out/r59_pytest.log
for v in 2 3; do BCAST_DZT_MINB=$v timeout 200 python tools/dz_tangent_probe.py | sed "s/^{/{\"minb\": $v, /" >> gpurun_out/r59_dz_tangent.jsonl; done; cat gpurun_out/r59_dz_tangent.jsonl
BCAST_DZT_MINB=2 timeout 200 ncu --set full --clock-control none --import-source on -k regex:k_dz_tangent -s 3 -c 1 -f -o gpurun_out/r59_dz_tangent_full python tools/dz_tangent_probe.py 4096 1024 2 > gpurun_out/r59_ncu.log 2>&1; tail -n 2 gpurun_out/r59_ncu.log
