mkdir -p gpurun_out
N=$1
for ov in 1 0; do
BROADCAST_B200_STEP_OVERLAP=$ov timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2964$ov bench.py --gpus $N --steps 40 --warmup 5 --no-jacobian --no-e2e --no-cpu-baseline > gpurun_out/r2_22_bench_n${N}_ov$ov.json 2> gpurun_out/r2_22_bench_n${N}_ov$ov.err
python - <<PY
import json
d=json.loads(open('gpurun_out/r2_22_bench_n${N}_ov$ov.json').read().strip().splitlines()[-1])
print('overlap', $ov, d['ms_per_step'], d['roofline']['kernel_ms'], d['checksum']['res_bits_sum_i64'], d['config']['step_issue'])
PY
tail -2 gpurun_out/r2_22_bench_n${N}_ov$ov.err
done
