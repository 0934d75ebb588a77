mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -q -k "streamed" 2>&1 | tail -15
for cfg in "1 8" "1 16" "1 32" "0 8"; do set -- $cfg
BROADCAST_B200_E2E_ROWS=$1 BROADCAST_B200_E2E_SLABS=$2 timeout 600 python bench.py --steps 10 --warmup 3 --no-jacobian --no-cpu-baseline > gpurun_out/r2_28_bench_r$1_s$2.json 2> gpurun_out/r2_28_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_28_bench_r$1_s$2.json').read().strip().splitlines()[-1])
print('rows $1 windows $2 e2e ms', d['e2e']['ms_per_step'], 'value', d['e2e']['value'], d['e2e'].get('result_matches_resident'), d['e2e']['h2d_bytes_per_step'])
PY
done
tail -2 gpurun_out/r2_28_bench.err
