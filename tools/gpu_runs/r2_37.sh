mkdir -p gpurun_out
timeout 600 python tools/csr_probe.py 4096x2048 2>&1 | tail -1 | tee gpurun_out/r2_37_csr_probe.jsonl
