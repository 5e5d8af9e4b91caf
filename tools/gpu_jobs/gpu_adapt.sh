#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gram_gpu.py tests/test_pca_gpu.py -x -q -m gpu 2>&1 | tail -4
python tools/eig_bench.py 2>&1 | tail -3
echo "--- adaptive on (default), int8"
SWEEP_CG=2 SWEEP_KBW=0,74 SWEEP_LEAD=0 SWEEP_REPS=8 python tools/sweep_gram.py 2>&1 | tail -2
echo "--- adaptive off, int8"
VPCA_ADAPTIVE=0 SWEEP_CG=2 SWEEP_KBW=0 SWEEP_LEAD=0 SWEEP_REPS=4 python tools/sweep_gram.py 2>&1 | tail -1
echo "--- adaptive on, e2m1 / bf16 / 5M variants int8"
SWEEP_DTYPE=e2m1 SWEEP_CG=2 SWEEP_KBW=0 SWEEP_LEAD=0 SWEEP_REPS=6 python tools/sweep_gram.py 2>&1 | tail -1
SWEEP_DTYPE=bf16 SWEEP_CG=2 SWEEP_KBW=0 SWEEP_LEAD=0 SWEEP_REPS=6 python tools/sweep_gram.py 2>&1 | tail -1
SWEEP_V=5000000 SWEEP_CG=2 SWEEP_KBW=0 SWEEP_LEAD=0 SWEEP_REPS=6 python tools/sweep_gram.py 2>&1 | tail -1
