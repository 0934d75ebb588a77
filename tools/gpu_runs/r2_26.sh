mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_parity_gpu.py -q -k "streamed" 2>&1 | tail -3
for cfg in "1 8" "0 8" "1 12" "1 16"; do set -- $cfg
BROADCAST_B200_E2E_TAPER=$1 BROADCAST_B200_E2E_SLABS=$2 timeout 600 python bench.py --steps 10 --warmup 3 --no-jacobian --no-cpu-baseline > gpurun_out/r2_26_bench_t$1_s$2.json 2> gpurun_out/r2_26_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_26_bench_t$1_s$2.json').read().strip().splitlines()[-1])
print('taper $1 slabs $2 e2e ms', d['e2e']['ms_per_step'], 'value', d['e2e']['value'])
PY
done
tail -2 gpurun_out/r2_26_bench.err
