timeout 900 python -m pytest tests/test_table_gpu.py tests/test_sweep_gpu.py tests/test_table_large_gpu.py -m gpu -x -q 2>&1 | tail -3
for v in 64 20 200; do
echo "== helper sleep $v"
[ $v != 64 ] && export STB_B200_LIB=$PWD/libstb_b200/lib/hs$v/libstb_b200.so
python tools/quick_time.py shape 200000 20000 0.7 1 2>&1 | tail -1
python tools/quick_time.py shape 200000 20000 0.7 3 2>&1 | tail -1
python tools/quick_sweep.py 2>&1 | tail -1
done
