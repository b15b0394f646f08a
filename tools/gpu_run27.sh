mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r01w_pytest_gpu.log 2>&1
tail -3 gpurun_out/r01w_pytest_gpu.log
timeout 200 python __graft_entry__.py smoke 2>&1 | tail -1
