#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_v3.json 2> gpurun_out/bench_v3.err
tail -c 5500 gpurun_out/bench_v3.json; tail -5 gpurun_out/bench_v3.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_v3.csv \
   python bench.py --steps 2 --warmup 1 --e2e-steps 0 --no-cpu-baseline --no-eig-check --no-alt > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log
ncu --set full --clock-control none --import-source on -k regex:gram_kernel -s 3 -c 1 -o gpurun_out/prof_gram_i8_v3 -f \
   python bench.py --steps 2 --warmup 1 --e2e-steps 0 --no-cpu-baseline --no-eig-check --no-alt > gpurun_out/ncu_full_i8.log 2>&1
tail -2 gpurun_out/ncu_full_i8.log
