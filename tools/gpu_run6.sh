mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r01g_pytest_gpu.log 2>&1
tail -5 gpurun_out/r01g_pytest_gpu.log
timeout 1200 python bench.py --workloads graph --graph-rows 12500000 --no-cpu-baseline --steps 3 --warmup 2 > gpurun_out/r01g_bench_graph_12m5.json 2> gpurun_out/r01g_bench_graph_12m5.err
tail -3 gpurun_out/r01g_bench_graph_12m5.err
ls -la gpurun_out
