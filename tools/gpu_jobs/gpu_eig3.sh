#!/bin/bash
timeout 600 python -m pytest tests/test_pca_gpu.py tests/test_driver_gpu.py -x -q -m gpu 2>&1 | tail -4
python tools/eig_bench.py 2>&1 | tail -3
EIG_N=300,700,1500,3000 python tools/eig_bench.py 2>&1 | tail -4
