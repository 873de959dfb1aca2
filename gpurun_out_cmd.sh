timeout 600 python tools/svd_check.py 2>&1 | tail -20
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_layers_gpu.py -q -m gpu -k "svd" 2>&1 | tail -5
