mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "streamed" 2>&1 | tail -3
for n in 4 8 16 32; do BROADCAST_B200_E2E_SLABS=$n timeout 600 python bench.py --no-jacobian --no-cpu-baseline --steps 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print($n, d['e2e']['ms_per_step'], d['e2e']['value'])" >> gpurun_out/r28_e2e.log; done; cat gpurun_out/r28_e2e.log
