#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pca_gpu.py tests/test_driver_gpu.py -x -q -m gpu 2>&1 | tail -6
echo "--- fused single-launch steps"
python tools/eig_bench.py 2>&1 | tail -3
echo "--- two-kernel steps"
VPCA_EIG_TWO_KERNELS=1 python tools/eig_bench.py 2>&1 | tail -3
timeout 300 python -m pytest tests/test_gram_gpu.py -x -q -m gpu -k "biobank" 2>&1 | tail -3
