mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r68_pytest.log
timeout 900 python bench.py > gpurun_out/r68_bench.json 2>gpurun_out/r68_bench.err; wc -l gpurun_out/r68_bench.json; tail -n 2 gpurun_out/r68_bench.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
