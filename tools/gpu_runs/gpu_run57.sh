mkdir -p gpurun_out
BROADCAST_B200_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_tangent_tile" -s 40 -c 2 -o gpurun_out/r57_tangent_tile_full python tools/jac_probe.py 2048x512 > gpurun_out/r57_ncu1.log 2>&1; tail -n 2 gpurun_out/r57_ncu1.log
BROADCAST_B200_NO_GRAPH=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_jac_assemble_rt" -s 1 -c 1 -o gpurun_out/r57_assemble_full python tools/jac_probe.py 2048x512 > gpurun_out/r57_ncu2.log 2>&1; tail -n 2 gpurun_out/r57_ncu2.log
