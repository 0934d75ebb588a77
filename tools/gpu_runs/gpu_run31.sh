set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r31_pytest.log; cat gpurun_out/r31_pytest.log
timeout 600 python bench.py --no-jacobian --no-cpu-baseline --no-e2e > gpurun_out/r31_bench.json 2>gpurun_out/r31_bench.err; cut -c1-250 gpurun_out/r31_bench.json; tail -n 3 gpurun_out/r31_bench.err
timeout 300 python tools/res_probe.py 8192x2048 2>&1 | cut -c1-100
