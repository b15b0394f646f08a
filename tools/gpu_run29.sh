mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_graph_gpu.py -m gpu -q -k "load_packed_index or beam_search_bit_exact" 2>&1 | tail -15
