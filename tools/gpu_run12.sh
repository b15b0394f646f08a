mkdir -p gpurun_out
set -x
timeout 300 python tools/text_latency.py > gpurun_out/r01m_text_latency.json 2> gpurun_out/r01m_text_latency.err
cat gpurun_out/r01m_text_latency.json; tail -2 gpurun_out/r01m_text_latency.err
timeout 900 python bench.py --workloads graph --graph-rows 12500000 --no-cpu-baseline --steps 3 --warmup 2 > gpurun_out/r01m_bench_graph_12m5.json 2> gpurun_out/r01m_bench_graph_12m5.err
tail -2 gpurun_out/r01m_bench_graph_12m5.err
