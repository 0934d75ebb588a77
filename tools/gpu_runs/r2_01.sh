set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
timeout 900 python -m pytest tests/test_parity_gpu.py -x -q -m gpu -k "boundary_fill_and_residual or variants_agree or nowall or c5 or C5 or full_size" 2>&1 | tail -15
timeout 600 python tools/res_probe.py 500x150 2048x512 8192x2048 > gpurun_out/r2_01_res_probe.jsonl 2> gpurun_out/r2_01_res_probe.err; cat gpurun_out/r2_01_res_probe.jsonl; tail -5 gpurun_out/r2_01_res_probe.err
