mkdir -p gpurun_out
N=$1
timeout 900 python -m pytest tests/test_multigpu_gpu.py -q 2>&1 | tail -4 | tee gpurun_out/r2_36_pytest.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29671 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_36_bench_n$N.json 2> gpurun_out/r2_36_bench_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_36_bench_n$N.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'ms', d['ms_per_step'], 'value', d['value'], 'checksum', d['checksum']['res_bits_sum_i64'], 'e2e ms', d['e2e']['ms_per_step'])
print(json.dumps(d['jacobian'])[:600])
PY
tail -2 gpurun_out/r2_36_bench_n$N.err
