mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r01h_pytest_gpu.log 2>&1
tail -4 gpurun_out/r01h_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r01h_smoke.log 2>&1
tail -2 gpurun_out/r01h_smoke.log
timeout 900 python bench.py --workloads graph --graph-data latent --no-cpu-baseline --steps 5 --warmup 3 > gpurun_out/r01h_bench_graph_latent.json 2> gpurun_out/r01h_bench_graph_latent.err
tail -2 gpurun_out/r01h_bench_graph_latent.err
ls -la gpurun_out
