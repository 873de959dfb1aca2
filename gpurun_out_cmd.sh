set -x
mkdir -p gpurun_out
timeout 300 python tools/svd_ab.py > gpurun_out/c15_svd_ab.log 2>&1
cat gpurun_out/c15_svd_ab.log
