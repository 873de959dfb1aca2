set -x
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_conv_gpu.py tests/test_layers_gpu.py -q -m gpu -x ) > gpurun_out/c4_pytest.log 2>&1
tail -4 gpurun_out/c4_pytest.log
( SDNQ_B200_BN=192 timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_layers_gpu.py -q -m gpu -x -k "mm or gemm or forward or layer or epilogue" ) > gpurun_out/c4_pytest_bn192.log 2>&1
tail -4 gpurun_out/c4_pytest_bn192.log
timeout 300 python tools/actq_ab.py > gpurun_out/c4_actq_tc.log 2>&1
SDNQ_B200_HADAMARD_BUTTERFLY=1 timeout 300 python tools/actq_ab.py > gpurun_out/c4_actq_bf.log 2>&1
paste -d'\n' gpurun_out/c4_actq_tc.log gpurun_out/c4_actq_bf.log | grep "hadamard=[1-9]" | head -24
grep "hadamard=0" gpurun_out/c4_actq_tc.log | head -8
timeout 300 ncu --set full --clock-control none --import-source on -k regex:act_quant -c 1 -s 2 -f -o gpurun_out/c4_actq_had_tc python tools/run_one.py actq 16384 3072 256 fp8 > gpurun_out/c4_ncu.log 2>&1
timeout 300 python tools/shape_breakdown.py sdxl > gpurun_out/c4_bd_sdxl.log 2>&1
tail -12 gpurun_out/c4_bd_sdxl.log
SDNQ_B200_BN=256 timeout 300 python tools/shape_breakdown.py sdxl > gpurun_out/c4_bd_sdxl_bn256.log 2>&1
tail -12 gpurun_out/c4_bd_sdxl_bn256.log
