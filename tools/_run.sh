timeout 900 python -m pytest tests/test_table_gpu.py tests/test_sweep_gpu.py -m gpu -x -q 2>&1 | tail -3
for w in 1 2 3 4; do for s in 1 2; do
echo "== weight $w spread $s"
STB_STRIP_WEIGHT=$w STB_STRIP_SPREAD=$s python tools/quick_time.py shape 200000 20000 0.7 1 2>&1 | tail -1
done; done
for w in 1 2 3; do
STB_STRIP_WEIGHT=$w python tools/quick_time.py shape 200000 20000 0.7 3 2>&1 | tail -1
STB_STRIP_WEIGHT=$w python tools/quick_sweep.py 2>&1 | tail -1
done
