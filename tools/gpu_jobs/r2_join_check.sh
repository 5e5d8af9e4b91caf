#!/bin/bash
# round 2: the device join tests after they were extended to compare with the oracle's restatement
timeout 600 python -m pytest tests/test_join_gpu.py tests/test_driver_gpu.py -q -m gpu --maxfail=3 --tb=short 2>&1 | tail -6 | cut -c1-300
