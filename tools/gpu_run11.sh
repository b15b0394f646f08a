mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests/test_graph_gpu.py -m gpu -q > gpurun_out/r01l_pytest_graph.log 2>&1
tail -6 gpurun_out/r01l_pytest_graph.log
