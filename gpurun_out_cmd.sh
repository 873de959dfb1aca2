mkdir -p gpurun_out
timeout 300 python tools/flat_tune.py > gpurun_out/c17_flat_tune.log 2>&1
cat gpurun_out/c17_flat_tune.log
