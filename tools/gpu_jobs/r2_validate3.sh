#!/bin/bash
# round 2, call 3: what breaks the pool / same-process-peer tests (full traceback of the first failure), the persistent
# Lanczos kernel (tests + timing), square tiling back as the default
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pool_gpu.py -x -q -m gpu --tb=long > gpurun_out/r2_pool_tests.log 2>&1
echo "pool tests rc=$?"; grep -n "Error\|error:\|vpca\|FAILED\|passed\|failed" gpurun_out/r2_pool_tests.log | head -40
timeout 900 python -m pytest tests/test_pca_gpu.py tests/test_parity_fullsize_gpu.py tests/test_driver_gpu.py -q -m gpu --maxfail=5 2>&1 | tail -25
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== eig bench: persistent Lanczos vs the five-kernel graph ==="
EIG_N=1092,2504,4096 EIG_MODES=auto EIG_REPS=7 timeout 300 python tools/eig_bench.py 2>&1 | tail -4
VPCA_LZ_PERSIST=0 EIG_N=2504 EIG_MODES=auto EIG_REPS=7 timeout 300 python tools/eig_bench.py 2>&1 | tail -2
