mkdir -p gpurun_out
( timeout 900 python -m pytest tests -q -m gpu ) > gpurun_out/c22_pytest.log 2>&1
tail -3 gpurun_out/c22_pytest.log
timeout 300 python tools/gemv_bench.py > gpurun_out/c22_gemv.log 2>&1
cat gpurun_out/c22_gemv.log
timeout 300 python tools/actq_ab.py 2>&1 | grep -E "M=16384 K=3072|M=18432" | grep -E "hadamard=256|hadamard=0" | head -8
