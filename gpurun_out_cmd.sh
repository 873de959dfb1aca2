set -x
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "rotated" ) > gpurun_out/c10_pytest.log 2>&1
tail -3 gpurun_out/c10_pytest.log
timeout 300 python tools/dequant_bw.py --rot > gpurun_out/c10_dequant_rot.log 2>&1
cat gpurun_out/c10_dequant_rot.log
SDNQ_B200_HADAMARD_BUTTERFLY=1 timeout 300 python tools/dequant_bw.py --rot > gpurun_out/c10_dequant_rot_bf.log 2>&1
cat gpurun_out/c10_dequant_rot_bf.log
timeout 300 python tools/conv_bench.py > gpurun_out/c10_conv_bench.log 2>&1
cat gpurun_out/c10_conv_bench.log
