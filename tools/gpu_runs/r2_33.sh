mkdir -p gpurun_out
timeout 600 python tools/newton_probe.py 500x150 0.2 lu 2>&1 | tail -1 | tee gpurun_out/r2_33_newton.jsonl
timeout 300 python tools/newton_probe.py 500x150 1.0 2>&1 | tail -1 | tee -a gpurun_out/r2_33_newton.jsonl
