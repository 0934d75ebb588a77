mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for n in 8 16; do BROADCAST_B200_E2E_SLABS=$n timeout 600 python bench.py --no-jacobian --no-cpu-baseline --steps 10 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print($n, 'e2e', d['e2e']['ms_per_step'], d['e2e']['value'], 'step', d['ms_per_step'])" >> gpurun_out/r35_e2e.log; done; cat gpurun_out/r35_e2e.log
