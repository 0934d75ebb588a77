mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_jac_assemble_rt -c 1 -o gpurun_out/r2_30_jac python tools/jac_probe.py 2048x512 > gpurun_out/r2_30_ncu.log 2>&1
tail -3 gpurun_out/r2_30_ncu.log
