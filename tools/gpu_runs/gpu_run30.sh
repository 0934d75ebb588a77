mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "jacobian or face_lin or csr" 2>&1 | tail -3
for cfg in 0 1 2 3 4; do BROADCAST_B200_JAC_CFG=$cfg timeout 300 python tools/jac_probe.py 2048x512 4096x1024 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l); print('cfg=$cfg', d['im'], d['jm'], 'interior_ms %.3f' % d['interior_ms'])
" >> gpurun_out/r30_jac_cfg.log; done; cat gpurun_out/r30_jac_cfg.log
