set -x
python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; tail -c 400 gpurun_out/bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_launch.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:fill_strip -s 4 -c 1 --csv --log-file gpurun_out/traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_traffic.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:fill_strip -s 1 -c 1 -o gpurun_out/prof_r1d -f python tools/prof_fill.py 40000 20000 0.7 1 2 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
