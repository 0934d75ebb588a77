mkdir -p gpurun_out
timeout 900 python tools/config_probe.py > gpurun_out/r39_configs.jsonl 2> gpurun_out/r39_configs.err; cat gpurun_out/r39_configs.jsonl; tail -n 3 gpurun_out/r39_configs.err
