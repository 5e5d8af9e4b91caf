#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
tail -c 7000 gpurun_out/bench_final.json; tail -5 gpurun_out/bench_final.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_final.csv \
   python bench.py --steps 2 --warmup 1 --e2e-steps 0 --no-cpu-baseline --no-eig-check --no-alt > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log
