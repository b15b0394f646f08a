mkdir -p gpurun_out
set -x
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_|k_gemm|k_mha|skinny" -s 194 -c 194 --csv --log-file gpurun_out/r01o_text_launches.csv python tools/text_latency.py > gpurun_out/r01o_ncu_text.log 2>&1
tail -3 gpurun_out/r01o_ncu_text.log
