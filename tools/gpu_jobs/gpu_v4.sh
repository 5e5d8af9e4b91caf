#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -6
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_v4.json 2> gpurun_out/bench_v4.err
tail -c 4000 gpurun_out/bench_v4.json | cut -c1-4000; tail -5 gpurun_out/bench_v4.err
python bench.py --steps 20 --warmup 3 --dtype e2m1 --no-cpu-baseline --e2e-steps 2 > gpurun_out/bench_v4_e2m1.json 2> gpurun_out/bench_v4_e2m1.err
tail -c 2500 gpurun_out/bench_v4_e2m1.json; tail -5 gpurun_out/bench_v4_e2m1.err
ncu --set full --clock-control none --import-source on -k regex:gram_kernel -s 3 -c 1 -o gpurun_out/prof_gram_mxf4_v4 -f \
   python bench.py --steps 2 --warmup 1 --e2e-steps 0 --no-cpu-baseline --no-eig-check --dtype e2m1 > gpurun_out/ncu_full_mxf4.log 2>&1
tail -2 gpurun_out/ncu_full_mxf4.log
echo "--- large N sweeps (stream-K path)"
SWEEP_N=10000 SWEEP_V=200000 SWEEP_PANEL=4096 SWEEP_DTYPE=bf16 SWEEP_CG=2 SWEEP_KBW=0 SWEEP_REPS=3 timeout 300 python tools/sweep_gram.py 2>&1 | tail -1 | cut -c1-420
SWEEP_N=10000 SWEEP_V=200000 SWEEP_PANEL=4096 SWEEP_CG=2 SWEEP_KBW=0 SWEEP_REPS=3 timeout 300 python tools/sweep_gram.py 2>&1 | tail -1 | cut -c1-420
