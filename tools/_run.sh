timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tools/quick_time.py shape 200000 20000 0.7 1 2>&1 | tail -1
python tools/quick_time.py shape 200000 20000 0.7 3 2>&1 | tail -1
python tools/quick_time.py shape 10000 1000 0.5 3 2>&1 | tail -1
python tools/quick_sweep.py 2>&1 | tail -1
python tools/quick_sweep.py 5001 167 1036 2>&1 | tail -1
