mkdir -p gpurun_out
timeout 600 python tools/slab_jac_probe.py 8192x2048 > gpurun_out/r41_slab.log 2>&1; cat gpurun_out/r41_slab.log
