mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r01q_pytest_gpu.log 2>&1
tail -4 gpurun_out/r01q_pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/r01q_smoke.log 2>&1
tail -1 gpurun_out/r01q_smoke.log
timeout 900 python bench.py > gpurun_out/r01q_bench.json 2> gpurun_out/r01q_bench.err
tail -2 gpurun_out/r01q_bench.err
timeout 400 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r01q_bench_ref.json 2>> gpurun_out/r01q_bench.err
