mkdir -p gpurun_out
nvidia-smi -L | tee gpurun_out/r2_06_gpus.log
timeout 600 python -m pytest tests/test_multigpu_gpu.py -x -q 2>&1 | tail -15 | tee gpurun_out/r2_06_pytest.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 20 --warmup 3 --no-jacobian > gpurun_out/r2_06_bench_n2.json 2> gpurun_out/r2_06_bench_n2.err
tail -c 2500 gpurun_out/r2_06_bench_n2.json; tail -5 gpurun_out/r2_06_bench_n2.err
BROADCAST_B200_HALO=nccl BROADCAST_B200_STEP_GRAPH=0 timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --steps 20 --warmup 3 --no-jacobian --no-e2e --no-cpu-baseline > gpurun_out/r2_06_bench_n2_nccl.json 2> gpurun_out/r2_06_bench_n2_nccl.err
tail -c 1200 gpurun_out/r2_06_bench_n2_nccl.json
