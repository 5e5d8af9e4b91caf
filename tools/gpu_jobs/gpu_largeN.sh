#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gram_gpu.py -x -q -m gpu -k "large_n or stream_k or small_dense or c1_calls" 2>&1 | tail -4
echo "--- large N sweeps (wave schedule)"
SWEEP_N=10000 SWEEP_V=200000 SWEEP_PANEL=4096 SWEEP_DTYPE=bf16 SWEEP_CG=2 SWEEP_KBW=0 SWEEP_REPS=3 timeout 300 python tools/sweep_gram.py 2>&1 | tail -1 | cut -c1-420
SWEEP_N=10000 SWEEP_V=200000 SWEEP_PANEL=4096 SWEEP_CG=2 SWEEP_KBW=0 SWEEP_REPS=3 timeout 300 python tools/sweep_gram.py 2>&1 | tail -1 | cut -c1-420
SWEEP_N=10000 SWEEP_V=200000 SWEEP_PANEL=4096 SWEEP_DTYPE=e2m1 SWEEP_CG=2 SWEEP_KBW=0 SWEEP_REPS=3 timeout 300 python tools/sweep_gram.py 2>&1 | tail -1 | cut -c1-420
SWEEP_N=30000 SWEEP_V=65536 SWEEP_PANEL=4096 SWEEP_CG=2 SWEEP_KBW=0 SWEEP_REPS=2 timeout 300 python tools/sweep_gram.py 2>&1 | tail -1 | cut -c1-420
