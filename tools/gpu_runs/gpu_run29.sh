set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r29_pytest.log; cat gpurun_out/r29_pytest.log
timeout 900 python bench.py > gpurun_out/r29_bench.json 2>gpurun_out/r29_bench.err; wc -l gpurun_out/r29_bench.json; tail -n 3 gpurun_out/r29_bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r29_bench_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r29_bench_launches.log 2>&1; tail -n 2 gpurun_out/r29_bench_launches.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r29_smoke.log 2>&1; tail -n 4 gpurun_out/r29_smoke.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r29_ref.json 2>/dev/null; cut -c1-300 gpurun_out/r29_ref.json
