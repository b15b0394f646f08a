mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_encoder_gpu.py tests/test_e2e_gpu.py -m gpu -q > gpurun_out/r01n_pytest_enc.log 2>&1
tail -6 gpurun_out/r01n_pytest_enc.log
timeout 300 python tools/text_latency.py > gpurun_out/r01n_text_latency.json 2> gpurun_out/r01n_text_latency.err
cat gpurun_out/r01n_text_latency.json; tail -2 gpurun_out/r01n_text_latency.err
