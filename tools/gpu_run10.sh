mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r01k_pytest_gpu.log 2>&1
tail -4 gpurun_out/r01k_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r01k_smoke.log 2>&1
tail -1 gpurun_out/r01k_smoke.log
timeout 600 python tools/graph_bench.py 1000000 64 > gpurun_out/r01k_graph_schedules_1m.json 2> gpurun_out/r01k_graph_schedules.err
tail -2 gpurun_out/r01k_graph_schedules.err
