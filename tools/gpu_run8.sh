mkdir -p gpurun_out
set -x
nproc
timeout 900 python bench.py > gpurun_out/r01i_bench.json 2> gpurun_out/r01i_bench.err
tail -3 gpurun_out/r01i_bench.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01i_bench_ref.json 2>> gpurun_out/r01i_bench.err
timeout 600 python bench.py --workloads e2e --no-cpu-baseline > gpurun_out/r01i_bench_e2e.json 2> gpurun_out/r01i_bench_e2e.err
tail -3 gpurun_out/r01i_bench_e2e.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_|k_mha|k_gemm" -c 1200 --csv --log-file gpurun_out/r01i_tower_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --workloads tower > gpurun_out/r01i_ncu_tower.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_greedy_search_wq" -s 64 -c 1 -o gpurun_out/r01i_greedy_1m_full python bench.py --workloads graph --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r01i_ncu_graph.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_|k_greedy|k_beam|k_prune|k_merge|k_apply|k_rabitq|k_flat|k_select|k_rerank" -s 400 -c 60 --csv --log-file gpurun_out/r01i_graph_launches.csv python bench.py --workloads graph --graph-rows 200000 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r01i_ncu_graph2.log 2>&1
ls -la gpurun_out
