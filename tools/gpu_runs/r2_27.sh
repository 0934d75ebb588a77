mkdir -p gpurun_out
timeout 300 python tools/e2e_timeline.py 8 0 2>&1 | tail -1 | tee gpurun_out/r2_27_timeline.jsonl
timeout 300 python tools/e2e_timeline.py 8 1 2>&1 | tail -1 | tee -a gpurun_out/r2_27_timeline.jsonl
