mkdir -p gpurun_out
BROADCAST_B200_RESIDUAL_L2DIST=0 timeout 300 compute-sanitizer --tool memcheck --print-limit 3 python tools/res_one.py 96x48 6 1 2>&1 | grep -v "^=========     \(Host\|    \)" | head -40 > gpurun_out/r2_08_memcheck_nopf.log
timeout 300 compute-sanitizer --tool memcheck --print-limit 3 python tools/res_one.py 96x48 6 1 2>&1 | grep -v "^=========     \(Host\|    \)" | head -40 > gpurun_out/r2_08_memcheck_pf.log
head -30 gpurun_out/r2_08_memcheck_nopf.log; echo ----; head -30 gpurun_out/r2_08_memcheck_pf.log
