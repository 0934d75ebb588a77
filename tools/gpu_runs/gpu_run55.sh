mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/r55_bench_n4.json 2> gpurun_out/r55_bench_n4.err; wc -l gpurun_out/r55_bench_n4.json; tail -n 3 gpurun_out/r55_bench_n4.err | cut -c1-300
