mkdir -p gpurun_out
set -x
timeout 600 python -m pytest tests/test_encoder_gpu.py -m gpu -q -k "full_depth" > gpurun_out/r01v_pytest_fulldepth.log 2>&1
tail -5 gpurun_out/r01v_pytest_fulldepth.log
