mkdir -p gpurun_out
timeout 600 python tools/csr_probe.py 2048x2048 1024x2048 > gpurun_out/r2_18_csr_probe.jsonl 2> gpurun_out/r2_18_csr_probe.err
cat gpurun_out/r2_18_csr_probe.jsonl; tail -3 gpurun_out/r2_18_csr_probe.err
