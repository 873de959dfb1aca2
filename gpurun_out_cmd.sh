set -x
mkdir -p gpurun_out
( timeout 600 python -m pytest tests/test_layers_gpu.py tests/test_conv_gpu.py -q -m gpu ) > gpurun_out/c16_pytest.log 2>&1
tail -4 gpurun_out/c16_pytest.log
timeout 600 python bench.py --workload sdxl_int4_svd_dequant --no-cpu-baseline > gpurun_out/c16_bench_int4_stream.json 2> gpurun_out/c16_bench_int4_stream.err
cut -c1-420 gpurun_out/c16_bench_int4_stream.json; tail -3 gpurun_out/c16_bench_int4_stream.err
SDNQ_B200_DEQUANT_STREAM=0 timeout 600 python bench.py --workload sdxl_int4_svd_dequant --no-cpu-baseline > gpurun_out/c16_bench_int4_nostream.json 2> gpurun_out/c16_bench_int4_nostream.err
cut -c1-420 gpurun_out/c16_bench_int4_nostream.json
