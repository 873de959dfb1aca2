set -x
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_layers_gpu.py tests/test_conv_gpu.py -q -m gpu ) > gpurun_out/c9_pytest.log 2>&1
tail -6 gpurun_out/c9_pytest.log
timeout 300 python tools/actq_ab.py > gpurun_out/c9_actq.log 2>&1
grep -E "hadamard=(0|256)" gpurun_out/c9_actq.log | head -20
timeout 300 python tools/shape_breakdown.py sdxl > gpurun_out/c9_bd_sdxl.log 2>&1
tail -10 gpurun_out/c9_bd_sdxl.log
timeout 300 ncu --set full --clock-control none --import-source on -k regex:act_quant -c 1 -s 2 -f -o gpurun_out/c9_actq_had_tc python tools/run_one.py actq 16384 3072 256 fp8 > gpurun_out/c9_ncu.log 2>&1
