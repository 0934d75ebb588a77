set -x
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_residual_tile -s 3 -c 1 -o gpurun_out/r1_residual_full python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r1_ncu_full.log 2>&1; tail -3 gpurun_out/r1_ncu_full.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1_jac_launches.csv python tools/jac_probe.py 2048x512 > gpurun_out/r1_jac_launches.log 2>&1; tail -2 gpurun_out/r1_jac_launches.log
timeout 600 ncu --set full --clock-control none -k regex:k_jac_block -s 13 -c 2 -o gpurun_out/r1_jacblock_full python tools/jac_probe.py 2048x512 > gpurun_out/r1_ncu_jac.log 2>&1; tail -3 gpurun_out/r1_ncu_jac.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r1_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r1_bench_launches.log 2>&1
ls -la gpurun_out
