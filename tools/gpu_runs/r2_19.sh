mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_strips_window_gpu.py tests/test_parity_gpu.py -q -k "csr or hybrid or banded or window" > gpurun_out/r2_19_pytest.log 2>&1
tail -5 gpurun_out/r2_19_pytest.log
timeout 600 python tools/csr_probe.py 1024x2048 > gpurun_out/r2_19_csr_probe.jsonl 2> gpurun_out/r2_19_csr_probe.err
cat gpurun_out/r2_19_csr_probe.jsonl; tail -3 gpurun_out/r2_19_csr_probe.err
