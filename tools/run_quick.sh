#!/bin/bash
# Development aid (GPU box): parity tests of the fill + timings of config 2 (S, S+V, V), the config-3 shape and config 1
mkdir -p gpurun_out
out=gpurun_out/quick.log
: > $out
if [ -z "$SKIP_TESTS" ]; then
  timeout 600 python -m pytest tests/test_table_gpu.py tests/test_sweep_gpu.py -x -q 2>&1 | tail -3 >> $out
fi
for fl in 1 3 2; do
  timeout 300 python tools/quick_time.py shape 200000 20000 0.7 $fl >> $out 2>&1
done
timeout 300 python tools/quick_time.py shape 50000 5000 0.7 1 >> $out 2>&1
timeout 300 python tools/quick_time.py shape 10000 1000 0.5 3 >> $out 2>&1
cat $out
