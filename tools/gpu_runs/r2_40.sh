mkdir -p gpurun_out
for cfg in -1 0 1 2 3; do for ch in 0 8; do
  if [ $cfg -ge 0 ]; then export BROADCAST_B200_STRIPS_CFG=$cfg; else unset BROADCAST_B200_STRIPS_CFG; fi
  if [ $ch -gt 0 ]; then export BROADCAST_B200_STRIP_CHAINS=$ch; else unset BROADCAST_B200_STRIP_CHAINS; fi
  echo "cfg $cfg chains $ch: $(timeout 200 python tools/strips_only.py 8192x2048 2>&1 | tail -1)"
done; done | tee gpurun_out/r2_40_strips_sweep.log
