set -x
mkdir -p gpurun_out
for N in 8 4; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/c14_bench_sdxl_int8_${N}gpu.json 2> gpurun_out/c14_bench_sdxl_${N}gpu.err
cut -c1-330 gpurun_out/c14_bench_sdxl_int8_${N}gpu.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2952$N bench.py --gpus $N --steps 5 --warmup 3 --workload flux_fp8 > gpurun_out/c14_bench_flux_fp8_${N}gpu.json 2> gpurun_out/c14_bench_flux_${N}gpu.err
cut -c1-330 gpurun_out/c14_bench_flux_fp8_${N}gpu.json
done
