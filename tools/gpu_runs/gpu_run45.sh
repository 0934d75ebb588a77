set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r45_pytest.log; cat gpurun_out/r45_pytest.log
timeout 900 python bench.py > gpurun_out/r45_bench.json 2>gpurun_out/r45_bench.err; wc -l gpurun_out/r45_bench.json; tail -n 3 gpurun_out/r45_bench.err
BROADCAST_B200_NO_GRAPH=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r45_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r45_bench_launches.log 2>&1; tail -n 2 gpurun_out/r45_bench_launches.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_residual_fast -s 3 -c 1 -o gpurun_out/r45_residual_full python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-jacobian > gpurun_out/r45_ncu_res.log 2>&1; tail -n 2 gpurun_out/r45_ncu_res.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r45_smoke.log 2>&1; tail -n 3 gpurun_out/r45_smoke.log
