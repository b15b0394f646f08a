mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests/test_graph_gpu.py tests/test_e2e_gpu.py -m gpu -q > gpurun_out/r01j_pytest_graph.log 2>&1
tail -4 gpurun_out/r01j_pytest_graph.log
timeout 600 python bench.py --workloads graph --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r01j_bench_graph.json 2> gpurun_out/r01j_bench_graph.err
tail -2 gpurun_out/r01j_bench_graph.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_greedy|k_prune|k_merge|k_apply|k_random|k_centroid|k_medioid|k_argmax" -c 330 --csv --log-file gpurun_out/r01j_build_launches.csv python bench.py --workloads graph --graph-rows 200000 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r01j_ncu_build.log 2>&1
ls -la gpurun_out | grep r01j
