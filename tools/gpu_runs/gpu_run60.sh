mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/r60_pytest.log
timeout 200 python tools/dz_tangent_probe.py > gpurun_out/r60_dz_tangent.jsonl 2> gpurun_out/r60_dz_tangent.err; cat gpurun_out/r60_dz_tangent.jsonl
timeout 900 python bench.py > gpurun_out/r60_bench.json 2>gpurun_out/r60_bench.err; wc -l gpurun_out/r60_bench.json; tail -n 3 gpurun_out/r60_bench.err
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
