set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/r15_pytest.log; cat gpurun_out/r15_pytest.log
timeout 300 python tools/res_probe.py 2048x512 8192x2048 > gpurun_out/r15_res_probe.log 2>&1; cat gpurun_out/r15_res_probe.log
timeout 900 python bench.py > gpurun_out/r15_bench.json 2>gpurun_out/r15_bench.err; cat gpurun_out/r15_bench.json; tail -n 5 gpurun_out/r15_bench.err
ls -la gpurun_out
