#!/bin/bash
# round 2: persistent Lanczos with the basis columns mirrored in shared memory, finer phase profile; eigensolver tests
mkdir -p gpurun_out
VPCA_LZ_PROF=1 EIG_N=2504 EIG_MODES=auto EIG_REPS=7 timeout 300 python tools/eig_bench.py 2>&1 | tail -2
VPCA_LZ_PROF=1 EIG_N=1092,4096 EIG_MODES=auto EIG_REPS=5 timeout 300 python tools/eig_bench.py 2>&1 | tail -4
timeout 600 python -m pytest tests/test_pca_gpu.py -q -m gpu --maxfail=5 --tb=short 2>&1 | tail -5 | cut -c1-300
