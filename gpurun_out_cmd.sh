mkdir -p gpurun_out
( timeout 900 python -m pytest tests -q -m gpu ) > gpurun_out/c20_pytest.log 2>&1
tail -3 gpurun_out/c20_pytest.log
timeout 300 python tools/dq_grid_tune.py 2>&1 | grep "float8_e4m3fn g-1 had0"
