mkdir -p gpurun_out
set -x
timeout 400 python tools/flat_latency.py > gpurun_out/r01r_flat_latency.json 2> gpurun_out/r01r_flat_latency.err
cat gpurun_out/r01r_flat_latency.json; tail -2 gpurun_out/r01r_flat_latency.err
