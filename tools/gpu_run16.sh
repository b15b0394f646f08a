mkdir -p gpurun_out
set -x
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"skinny" -s 220 -c 4 -o gpurun_out/r01p_skinny_full python tools/text_latency.py > gpurun_out/r01p_ncu_skinny.log 2>&1
tail -3 gpurun_out/r01p_ncu_skinny.log
