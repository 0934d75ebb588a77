set -x
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r6_bench.json 2>gpurun_out/r6_bench.err; cat gpurun_out/r6_bench.json; tail -5 gpurun_out/r6_bench.err
nvidia-smi --query-gpu=memory.used,memory.total --format=csv
