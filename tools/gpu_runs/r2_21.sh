mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu -x > gpurun_out/r2_21_pytest.log 2>&1
tail -5 gpurun_out/r2_21_pytest.log
timeout 1200 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2_21_bench.json 2> gpurun_out/r2_21_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_21_bench.json').read().strip().splitlines()[-1])
print(json.dumps(d.get('jacobian'))[:1500]); print(d['ms_per_step'], d['roofline']['frac'])
PY
tail -3 gpurun_out/r2_21_bench.err
