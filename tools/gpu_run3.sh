mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r01d_pytest_gpu.log
timeout 900 python bench.py --workloads graph --steps 5 --warmup 3 > gpurun_out/r01d_bench_graph.json 2> gpurun_out/r01d_bench_graph.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_greedy_search_wq|k_beam_search_wq" -s 51 -c 7 -o gpurun_out/r01d_graph_full python bench.py --workloads graph --graph-rows 200000 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r01d_ncu_graph.log 2>&1
ls -la gpurun_out
