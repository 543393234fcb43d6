timeout 900 python -m pytest tests/test_samplers_gpu.py tests/test_sweep_gpu.py -m gpu -x -q 2>&1 | tail -3
python tools/_t.py 2>&1 | tail -3
