mkdir -p gpurun_out
timeout 200 python -m pytest tests -m gpu -x -q -k "jacobian_with_isothermal or c_host or c_abi_context or dz_tangent_colour" 2>&1 | tail -8 | tee gpurun_out/r71_pytest.log
