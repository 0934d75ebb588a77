set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/r18_bench_n8.json 2> gpurun_out/r18_bench_n8.err; cat gpurun_out/r18_bench_n8.json; tail -n 5 gpurun_out/r18_bench_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r18_bench_n4.json 2> gpurun_out/r18_bench_n4.err; cat gpurun_out/r18_bench_n4.json; tail -n 5 gpurun_out/r18_bench_n4.err
