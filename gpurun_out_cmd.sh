set -x
timeout 300 python -m pytest tests/test_kernels_gpu.py -x -q -m gpu -k "flat or dequant" 2>&1 | tail -8
timeout 300 python tools/dequant_bw.py > gpurun_out/dequant_bw.log 2>&1
cat gpurun_out/dequant_bw.log
