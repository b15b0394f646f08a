set -x
mkdir -p gpurun_out
nproc; nvidia-smi -L
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r01b_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r01b_bench.json 2> gpurun_out/r01b_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01b_bench_ref.json 2>> gpurun_out/r01b_bench.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r01b_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --workloads tower > gpurun_out/r01b_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_gemm_tn|k_mha_tc|k_layernorm" -s 300 -c 9 -o gpurun_out/r01b_tower_full python bench.py --steps 1 --warmup 1 --no-cpu-baseline --workloads tower > gpurun_out/r01b_ncu_full.log 2>&1
ls -la gpurun_out
