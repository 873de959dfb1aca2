set -x
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c13_smoke.log 2>&1
cat gpurun_out/c13_smoke.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/c13_bench_sdxl_int8_2gpu.json 2> gpurun_out/c13_bench_2gpu.err
cat gpurun_out/c13_bench_sdxl_int8_2gpu.json | cut -c1-600
tail -3 gpurun_out/c13_bench_2gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 5 --warmup 3 --workload flux_fp8 > gpurun_out/c13_bench_flux_fp8_2gpu.json 2> gpurun_out/c13_bench_flux_2gpu.err
cat gpurun_out/c13_bench_flux_fp8_2gpu.json | cut -c1-600
