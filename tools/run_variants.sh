#!/bin/bash
# Development aid (GPU box): times config 2 (and the config-3 shape) with every variant library under libstb_b200/lib/exp/.
mkdir -p gpurun_out
out=gpurun_out/variants.log
: > $out
for so in libstb_b200/lib/exp/libstb_b200_*.so; do
  name=$(basename $so .so); name=${name#libstb_b200_}
  echo "== $name" >> $out
  STB_B200_LIB=$PWD/$so timeout 300 python tools/quick_time.py shape 200000 20000 0.7 1 >> $out 2>&1
  STB_B200_LIB=$PWD/$so timeout 300 python tools/quick_time.py shape 50000 5000 0.7 1 >> $out 2>&1
  if [ -n "$VARIANTS_SV" ]; then
    STB_B200_LIB=$PWD/$so timeout 300 python tools/quick_time.py shape 200000 20000 0.7 3 >> $out 2>&1
    STB_B200_LIB=$PWD/$so timeout 300 python tools/quick_time.py shape 200000 20000 0.7 2 >> $out 2>&1
  fi
done
cat $out
