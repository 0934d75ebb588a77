mkdir -p gpurun_out
for n in 8 16 32; do BROADCAST_B200_E2E_SLABS=$n timeout 600 python bench.py --no-jacobian --no-cpu-baseline --steps 5 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print($n, d['e2e']['ms_per_step'], d['e2e']['value'])" >> gpurun_out/r23_e2e.log; done; cat gpurun_out/r23_e2e.log
