set -x
mkdir -p gpurun_out
rm -f gpurun_out/r2_parity_configs.jsonl
python tools/res_one.py 8192x2048 0 10
python tools/res_one.py 8192x2048 4 10
timeout 1500 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "variants_agree or full_size" 2>&1 | tail -5
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_residual_march -s 2 -c 1 -o gpurun_out/r2_04_march python tools/res_one.py 8192x2048 0 2 > gpurun_out/r2_04_ncu.log 2>&1; tail -3 gpurun_out/r2_04_ncu.log
timeout 2400 python -m pytest tests/test_configs_gpu.py -x -q -m gpu 2>&1 | tail -15
cat gpurun_out/r2_parity_configs.jsonl
