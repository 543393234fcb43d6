python -m pytest tests/test_table_gpu.py tests/test_sweep_gpu.py tests/test_fuzz_gpu.py -x -q 2>&1 | tail -2
for M in 160 1600 20000; do python tools/quick_time.py shape 200000 $M 0.7 1; done
python tools/quick_time.py shape 200000 20000 0.7 3
python tools/quick_time.py shape 200000 20000 0.7 2
python tools/quick_time.py shape 50000 5000 0.7 1
python tools/quick_time.py shape 10000 1000 0.5 3
python tools/quick_sweep.py 50000 5000 64 | tail -1
