mkdir -p gpurun_out
for rep in 1 2; do for es in 1 0; do BROADCAST_B200_RESIDUAL_EARLY_SENSOR=$es timeout 300 python tools/res_probe.py 8192x2048 2>&1 | cut -c1-80 | sed "s/^/early=$es /"; done; done
