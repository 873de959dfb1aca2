mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_conv_gpu.py -q -m gpu ) > gpurun_out/c23_pytest.log 2>&1
tail -12 gpurun_out/c23_pytest.log
timeout 300 python tools/conv_bench.py > gpurun_out/c23_conv_bench.log 2>&1
cat gpurun_out/c23_conv_bench.log
