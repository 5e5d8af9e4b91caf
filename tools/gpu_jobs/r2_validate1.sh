#!/bin/bash
# round 2, call 1: the refactored boundary (lanes, pool, local peers, band-only Grams, C harnesses) + a bench sanity run
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -15
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_v1.json 2> gpurun_out/r2_bench_v1.err
tail -c 3000 gpurun_out/r2_bench_v1.json; tail -5 gpurun_out/r2_bench_v1.err
