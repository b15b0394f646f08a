mkdir -p gpurun_out
set -x
nvidia-smi -L
timeout 600 python -m pytest tests/test_graph_gpu.py -m gpu -q -k "generate_index_shard" > gpurun_out/r01f_pytest_shard.log 2>&1
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r01f_bench_n2.json 2> gpurun_out/r01f_bench_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/r01f_bench_n2_ref.json 2>> gpurun_out/r01f_bench_n2.err
tail -5 gpurun_out/r01f_bench_n2.err
ls -la gpurun_out
