#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_plink.py tests/test_gram_gpu.py tests/test_driver_gpu.py -x -q -m gpu 2>&1 | tail -8
