set -x
mkdir -p gpurun_out
for seg in 64 128 256 512 2048; do BROADCAST_B200_MARCH_SEG=$seg python tools/res_one.py 8192x2048 0 10; done 2>&1 | grep variant
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_residual_march -s 2 -c 1 -o gpurun_out/r2_02_march python tools/res_one.py 8192x2048 0 2 > gpurun_out/r2_02_ncu.log 2>&1; tail -3 gpurun_out/r2_02_ncu.log
