mkdir -p gpurun_out
timeout 300 python tools/pcie_probe.py > gpurun_out/r22_pcie.log 2>&1; cat gpurun_out/r22_pcie.log
