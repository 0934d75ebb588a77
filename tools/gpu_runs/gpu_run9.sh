set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r9_bench_n2.json 2> gpurun_out/r9_bench_n2.err; cat gpurun_out/r9_bench_n2.json; tail -5 gpurun_out/r9_bench_n2.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r9_ref.json 2> gpurun_out/r9_ref.err; cat gpurun_out/r9_ref.json; tail -3 gpurun_out/r9_ref.err
