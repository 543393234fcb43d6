tools/ubench_steps 2>&1 | head -16
timeout 900 python -m pytest tests/test_table_gpu.py tests/test_sweep_gpu.py -m gpu -x -q 2>&1 | tail -3
for k in 6 5; do
echo "== K=$k"
STB_STRIP_K=$k python tools/quick_time.py shape 200000 20000 0.7 1 2>&1 | tail -1
STB_STRIP_K=$k python tools/quick_time.py shape 200000 20000 0.7 3 2>&1 | tail -1
STB_STRIP_K=$k STB_STRIP_SPREAD=0 python tools/quick_time.py shape 200000 20000 0.7 1 2>&1 | tail -1
done
python tools/quick_sweep.py 2>&1 | tail -1
