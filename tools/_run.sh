timeout 900 python -m pytest tests/test_dropin_gpu.py -m gpu -x -q 2>&1 | tail -30
oracle/_ref/dropin_list -a 0.1 -N 100 -T 20 2>&1 | head -8
