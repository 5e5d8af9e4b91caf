#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_pca_gpu.py -x -q -m gpu 2>&1 | tail -4
EIG_N=1092,2504 EIG_REPS=3 timeout 100 python tools/eig_bench.py 2>&1 | tail -4 | tee gpurun_out/eig_bench_v14.jsonl
