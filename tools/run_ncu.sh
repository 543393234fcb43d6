#!/bin/bash
# Development aid (GPU box): ncu capture of one config-2 fill (source counters, stall sampling, memory and scheduler sections)
mkdir -p gpurun_out
tag=${1:-r2}
flags=${2:-1}
timeout 900 ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --section MemoryWorkloadAnalysis \
  --section ComputeWorkloadAnalysis --section SpeedOfLight --section LaunchStats --section Occupancy \
  --clock-control none --import-source on -k regex:fill_strip -s 1 -c 1 -o gpurun_out/$tag -f \
  python tools/prof_fill.py 200000 20000 0.7 $flags 2 > gpurun_out/ncu_$tag.log 2>&1
tail -3 gpurun_out/ncu_$tag.log
