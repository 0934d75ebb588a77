mkdir -p gpurun_out
N=$1
nvidia-smi -L | wc -l
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29641 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_16_bench_n$N.json 2> gpurun_out/r2_16_bench_n$N.err
tail -c 4000 gpurun_out/r2_16_bench_n$N.json; tail -4 gpurun_out/r2_16_bench_n$N.err
