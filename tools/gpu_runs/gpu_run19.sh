set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r19_pytest.log; cat gpurun_out/r19_pytest.log
timeout 300 python tools/res_probe.py 2048x512 8192x2048 > gpurun_out/r19_res_probe.log 2>&1; cat gpurun_out/r19_res_probe.log
timeout 600 python bench.py --no-jacobian --no-cpu-baseline > gpurun_out/r19_bench.json 2>gpurun_out/r19_bench.err; wc -l gpurun_out/r19_bench.json; cut -c1-700 gpurun_out/r19_bench.json; tail -n 3 gpurun_out/r19_bench.err
