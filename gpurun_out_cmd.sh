timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -x -k "packed_int4" 2>&1 | tail -6
timeout 600 python -m pytest tests/test_layers_gpu.py -q -m gpu -k "rowwise or forward_matches" 2>&1 | tail -4
