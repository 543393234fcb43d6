timeout 900 python -m pytest tests/test_table_gpu.py tests/test_table_large_gpu.py tests/test_sweep_gpu.py -m gpu -x -q 2>&1 | tail -15
python tools/quick_time.py shape 200000 20000 0.7 1 2>&1 | tail -1
python tools/quick_time.py shape 100000 40000 0.7 1 2>&1 | tail -1
