set -x
mkdir -p gpurun_out
for d in 0 296 592 1184 2368; do BROADCAST_B200_RESIDUAL_L2DIST=$d timeout 300 python tools/res_probe.py 8192x2048 2>&1 | cut -c1-120 | sed "s/^/dist=$d /" >> gpurun_out/r21_l2dist.log; done; cat gpurun_out/r21_l2dist.log
timeout 900 python -m pytest tests -m gpu -x -q -k "residual or slab or streamed" 2>&1 | tail -5 > gpurun_out/r21_pytest.log; cat gpurun_out/r21_pytest.log
