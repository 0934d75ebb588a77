mkdir -p gpurun_out
N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29681 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_39_bench_n$N.json 2> gpurun_out/r2_39_bench_n$N.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_39_bench_n$N.json').read().strip().splitlines()[-1])
print('N', d['n_gpus'], 'ms', d['ms_per_step'], 'value', d['value'], 'checksum', d['checksum']['res_bits_sum_i64'], 'e2e ms', d['e2e']['ms_per_step'])
j=d['jacobian']; print({k:j.get(k) for k in ('assembly_s','interior_blocks_ms','strips_and_fill_ms','csr_ms','assembly_plus_csr_s','gather_s')})
PY
tail -2 gpurun_out/r2_39_bench_n$N.err
