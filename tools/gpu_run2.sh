mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r01c_pytest_gpu.log
timeout 600 python tools/graph_bench.py 200000 64 > gpurun_out/r01c_graph_modes_200k.json 2> gpurun_out/r01c_graph_modes.err
timeout 900 python bench.py --workloads graph --steps 5 --warmup 3 > gpurun_out/r01c_bench_graph.json 2> gpurun_out/r01c_bench_graph.err
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_greedy_search_wq|k_beam_search" -c 3 -o gpurun_out/r01c_graph_full python bench.py --workloads graph --graph-rows 200000 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r01c_ncu_graph.log 2>&1
ls -la gpurun_out
