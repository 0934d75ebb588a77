set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r1_pytest.log; tail -5 gpurun_out/r1_pytest.log
timeout 600 python bench.py > gpurun_out/r1_bench.json 2> gpurun_out/r1_bench.err; cat gpurun_out/r1_bench.json; tail -3 gpurun_out/r1_bench.err
timeout 900 python tools/jac_probe.py 500x150 630x300 2048x512 > gpurun_out/r1_jac_probe.log 2>&1; cat gpurun_out/r1_jac_probe.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_residual_tile -s 3 -c 1 -o gpurun_out/r1_residual_full python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r1_ncu_full.log 2>&1; tail -3 gpurun_out/r1_ncu_full.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_jac_block -c 29 -o gpurun_out/r1_jacblock_full python tools/jac_probe.py 1024x256 > gpurun_out/r1_ncu_jac.log 2>&1; tail -3 gpurun_out/r1_ncu_jac.log
ls -la gpurun_out
