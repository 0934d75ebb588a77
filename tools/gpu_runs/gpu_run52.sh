mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/r52_bench.json 2>gpurun_out/r52_bench.err; wc -l gpurun_out/r52_bench.json; tail -n 3 gpurun_out/r52_bench.err
timeout 300 python tools/config_probe.py > gpurun_out/r52_configs.jsonl 2>/dev/null; cut -c1-400 gpurun_out/r52_configs.jsonl
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2
