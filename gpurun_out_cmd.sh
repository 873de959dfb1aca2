mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_layers_gpu.py -q -m gpu -x -k "not mm" 2>&1 | tail -4
python tools/bw_check.py 2>&1 | grep -v "^copy"
