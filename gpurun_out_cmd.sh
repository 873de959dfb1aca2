mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_kernels_gpu.py -q -m gpu -k "mm" 2>&1 | tail -5
timeout 900 python tools/bench_kernels.py 2>&1 | tee gpurun_out/bench_kernels.log | grep -v dequant | python -c "
import sys, json
for l in sys.stdin:
    try:
        k, j = l.split(' ', 1); d = json.loads(j)
    except Exception:
        print(l.strip()); continue
    print(k, 'int8 %.3f ms %.0f TF (cublas %.0f) | fp8 %.0f TF (cublas %.0f)' % (d['ours_int8_ms'], d['ours_int8_tflops'], d.get('cublaslt_int8_tflops', 0), d['ours_fp8_tflops'], d.get('cublaslt_fp8_tflops', 0)))
"
