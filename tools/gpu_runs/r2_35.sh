mkdir -p gpurun_out
timeout 1800 python -m pytest tests -q -m gpu > gpurun_out/r2_35_pytest.log 2>&1
tail -6 gpurun_out/r2_35_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4 | tee gpurun_out/r2_35_smoke.log
timeout 1500 python bench.py > gpurun_out/r2_35_bench.json 2> gpurun_out/r2_35_bench.err
tail -c 5000 gpurun_out/r2_35_bench.json; tail -3 gpurun_out/r2_35_bench.err
