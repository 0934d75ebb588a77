mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "csr or jacobian" 2>&1 | tail -3
timeout 600 python tools/csr_probe.py 500x150 2048x512 > gpurun_out/r43_csr_probe.log 2>&1; cat gpurun_out/r43_csr_probe.log
for cfg in 0 5; do BROADCAST_B200_JAC_CFG=$cfg timeout 300 python tools/jac_probe.py 4096x1024 2>&1 | cut -c1-200 | sed "s/^/cfg=$cfg /"; done
