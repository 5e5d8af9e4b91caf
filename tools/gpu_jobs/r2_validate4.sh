#!/bin/bash
# round 2, call 4: same-process peers after the kernel preload (lazy module loading vs spinning barriers), the device join /
# merge (f-3), the persistent Lanczos with the parallel cross-block reduction + its phase profile
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_pool_gpu.py tests/test_join_gpu.py -q -m gpu --maxfail=4 --tb=short > gpurun_out/r2_pool_join_tests.log 2>&1
echo "pool+join tests rc=$?"; tail -15 gpurun_out/r2_pool_join_tests.log | cut -c1-250
timeout 600 python -m pytest tests/test_pca_gpu.py tests/test_driver_gpu.py -q -m gpu --maxfail=5 2>&1 | tail -8
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
echo "=== eig bench: persistent Lanczos (phase profile of block 0) ==="
VPCA_LZ_PROF=1 EIG_N=1092,2504,4096 EIG_MODES=auto EIG_REPS=7 timeout 300 python tools/eig_bench.py 2>&1 | tail -6
EIG_N=2504,10000 EIG_MODES=auto EIG_REPS=5 timeout 300 python tools/eig_bench.py 2>&1 | tail -2
