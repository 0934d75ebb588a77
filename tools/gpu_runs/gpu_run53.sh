mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests -m gpu -x -q -k "hybrid_jacobian or csr_kernels or variants or smallest or two_parts" > gpurun_out/r53_memcheck.log 2>&1; echo "exit $?" >> gpurun_out/r53_memcheck.log; grep -E "ERROR SUMMARY|Invalid|exit|passed|failed|out of bounds" gpurun_out/r53_memcheck.log | head -20
