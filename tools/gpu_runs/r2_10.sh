mkdir -p gpurun_out
for d in 15 7 11 13 14 12 0; do
  echo "== dbg $d"; BROADCAST_B200_RESIDUAL_L2DIST=0 BROADCAST_B200_BULK_DEBUG=$d timeout 120 python tools/res_one.py 96x48 6 1 2>&1 | tail -2
done 2>&1 | tee gpurun_out/r2_10_dbg.log
