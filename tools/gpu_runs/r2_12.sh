mkdir -p gpurun_out
for d in 11 27 43 75 139 59 203; do
  echo "== dbg $d"; BROADCAST_B200_RESIDUAL_L2DIST=0 BROADCAST_B200_BULK_DEBUG=$d timeout 120 python tools/res_one.py 96x48 6 1 2>&1 | tail -1
done 2>&1 | tee gpurun_out/r2_12_dbg.log
