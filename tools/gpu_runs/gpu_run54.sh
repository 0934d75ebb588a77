mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 7 --print-limit 20 python -m pytest tests -m gpu -x -q -k "variants or two_parts or csr_kernels or (hybrid_jacobian and bl)" > gpurun_out/r54_racecheck.log 2>&1; echo "exit $?" >> gpurun_out/r54_racecheck.log; grep -E "RACECHECK SUMMARY|hazard|exit|passed|failed|Error" gpurun_out/r54_racecheck.log | head -20
