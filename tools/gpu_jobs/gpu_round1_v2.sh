#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pca_gpu.py tests/test_driver_gpu.py -x -q -m gpu 2>&1 | tail -8
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_v2.json 2> gpurun_out/bench_v2.err
tail -c 5000 gpurun_out/bench_v2.json; tail -5 gpurun_out/bench_v2.err
python bench.py --steps 20 --warmup 3 --dtype e2m1 --no-cpu-baseline > gpurun_out/bench_v2_e2m1.json 2> gpurun_out/bench_v2_e2m1.err
tail -c 3500 gpurun_out/bench_v2_e2m1.json; tail -5 gpurun_out/bench_v2_e2m1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_v2.csv \
   python bench.py --steps 2 --warmup 1 --e2e-steps 0 --no-cpu-baseline --no-eig-check --no-alt > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log
ncu --set full --clock-control none --import-source on -k regex:gram_kernel -s 3 -c 1 -o gpurun_out/prof_gram_i8_v2 -f \
   python bench.py --steps 2 --warmup 1 --e2e-steps 0 --no-cpu-baseline --no-eig-check --no-alt > gpurun_out/ncu_full_i8.log 2>&1
tail -2 gpurun_out/ncu_full_i8.log
ncu --set full --clock-control none --import-source on -k regex:gram_kernel -s 3 -c 1 -o gpurun_out/prof_gram_e2m1_v2 -f \
   python bench.py --steps 2 --warmup 1 --e2e-steps 0 --no-cpu-baseline --no-eig-check --dtype e2m1 > gpurun_out/ncu_full_e2m1.log 2>&1
tail -2 gpurun_out/ncu_full_e2m1.log
ls -la gpurun_out | tail -12
