#!/bin/bash
# Development aid (GPU box): producer / consumer phase counters (-DSTB_PROFILE_PRODUCER builds) on config 2
mkdir -p gpurun_out
out=gpurun_out/prof.log
: > $out
for so in libstb_b200/lib/exp/libstb_b200_prof*.so; do
  echo "== $so" >> $out
  STB_PROFILE_PRINT=1 STB_B200_LIB=$PWD/$so timeout 300 python tools/prof_fill.py 200000 20000 0.7 1 1 >> $out 2>&1
done
cat $out
