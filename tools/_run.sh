timeout 900 python -m pytest tests/test_table_gpu.py tests/test_config5_gpu.py tests/test_dropin_gpu.py -m gpu -x -q 2>&1 | tail -5
