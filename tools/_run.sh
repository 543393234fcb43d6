for s in 0 1; do
echo "== spread $s"
STB_STRIP_SPREAD=$s python tools/quick_time.py shape 200000 20000 0.7 3 2>&1 | tail -1
STB_STRIP_SPREAD=$s python tools/quick_time.py shape 200000 20000 0.7 5 2>&1 | tail -1
STB_STRIP_SPREAD=$s python tools/quick_sweep.py 2>&1 | tail -3
done
