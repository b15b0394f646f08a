mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r01e_pytest_gpu.log
timeout 900 python bench.py --workloads graph --steps 5 --warmup 3 > gpurun_out/r01e_bench_graph.json 2> gpurun_out/r01e_bench_graph.err
ls -la gpurun_out
