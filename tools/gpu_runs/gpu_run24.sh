set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r24_pytest.log; cat gpurun_out/r24_pytest.log
for cfg in 0 1 2 3; do BROADCAST_B200_STRIPS_CFG=$cfg timeout 300 python tools/jac_probe.py 500x150 4096x1024 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('cfg=$cfg', d['im'], d['jm'], 'hybrid_ms %.2f interior_ms %.2f strips_ms %.2f' % (d['hybrid_ms'], d['interior_ms'], d['hybrid_ms'] - d['interior_ms']))
" >> gpurun_out/r24_strips_cfg.log; done; cat gpurun_out/r24_strips_cfg.log
