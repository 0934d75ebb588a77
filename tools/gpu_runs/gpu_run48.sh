mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/r48_bench.json 2>gpurun_out/r48_bench.err; wc -l gpurun_out/r48_bench.json; tail -n 3 gpurun_out/r48_bench.err
timeout 300 python tools/config_probe.py > gpurun_out/r48_configs.jsonl 2>/dev/null; cat gpurun_out/r48_configs.jsonl | cut -c1-420
