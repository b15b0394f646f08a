mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests/test_flat_gpu.py -m gpu -q > gpurun_out/r01s_pytest_flat.log 2>&1
tail -3 gpurun_out/r01s_pytest_flat.log
timeout 400 python tools/flat_latency.py > gpurun_out/r01s_flat_latency.json 2> gpurun_out/r01s_flat_latency.err
cat gpurun_out/r01s_flat_latency.json; tail -2 gpurun_out/r01s_flat_latency.err
