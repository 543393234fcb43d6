#!/bin/bash
# GPU box: the round-2 ncu evidence for profiles/ -- (1) the bench command plain, (2) its launch list, (3) one --set full
# capture of the fill kernel at the full config-2 size, (4) DRAM bytes of one launch.
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras"
$B > gpurun_out/r2_bench_plain.json 2> gpurun_out/r2_bench_plain.err || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench.csv \
  $B > gpurun_out/r2_ncu_launch.log 2>&1
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:fill_strip -s 4 -c 1 -o gpurun_out/r2_fill_strip -f \
  $B > gpurun_out/r2_ncu_full.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none \
  -k regex:fill_strip -s 4 -c 1 --csv --log-file gpurun_out/r2_traffic_bench.csv $B > gpurun_out/r2_ncu_traffic.log 2>&1
tail -2 gpurun_out/r2_ncu_full.log; tail -3 gpurun_out/r2_traffic_bench.csv | cut -c1-400
