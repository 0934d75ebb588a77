mkdir -p gpurun_out
rm -f gpurun_out/r61_e2e_slabs.jsonl
for n in 8 16 32 64; do BROADCAST_B200_E2E_SLABS=$n timeout 300 python bench.py --no-jacobian --no-cpu-baseline --steps 10 --warmup 3 2>/dev/null | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print(json.dumps({'slabs': $n, 'e2e_ms': d['e2e']['ms_per_step'], 'e2e_value': d['e2e']['value'], 'h2d': d['e2e']['h2d_bytes_per_step'], 'd2h': d['e2e']['d2h_bytes_per_step']}))" >> gpurun_out/r61_e2e_slabs.jsonl; done
cat gpurun_out/r61_e2e_slabs.jsonl
