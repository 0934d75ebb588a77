mkdir -p gpurun_out
for cfg in 5; do BROADCAST_B200_JAC_CFG=$cfg timeout 300 python tools/jac_probe.py 2048x512 8192x2048 2>&1 | grep interior_ms | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); print('cfg $cfg', d['im'], d['jm'], 'interior_ms', round(d['interior_ms'],3), 'hybrid_ms', round(d['hybrid_ms'],3))
"; done | tee gpurun_out/r2_32_jac.log
