set -x
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -q -m gpu ) > gpurun_out/c12_pytest.log 2>&1
tail -12 gpurun_out/c12_pytest.log
timeout 600 python bench.py --workload flux_fp8 --steps 5 --warmup 3 > gpurun_out/c12_bench_flux_fp8.json 2> gpurun_out/c12_bench_flux.err
cat gpurun_out/c12_bench_flux_fp8.json | cut -c1-400
timeout 600 python bench.py > gpurun_out/c12_bench_sdxl_int8.json 2> gpurun_out/c12_bench_sdxl.err
cat gpurun_out/c12_bench_sdxl_int8.json | cut -c1-400
