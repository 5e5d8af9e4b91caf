#!/bin/bash
mkdir -p gpurun_out
for t in test_encode_tile_e2m1 "test_gram_e2m1_dense_and_calls[1]" "test_gram_e2m1_dense_and_calls[2]" test_synth_device_e2m1_matches_oracle test_gram_e2m1_resident_matches_int8; do
  timeout 120 python -m pytest "tests/test_gram_gpu.py::$t" -x -q -m gpu > gpurun_out/e2m1_$$.log 2>&1; rc=$?
  echo "$t rc=$rc"; if [ $rc -ne 0 ]; then tail -25 gpurun_out/e2m1_$$.log; fi
done
echo "--- tx full variant"
VPCA_E2M1_TX_FULL=1 timeout 60 python -m pytest "tests/test_gram_gpu.py::test_gram_e2m1_dense_and_calls[2]" -x -q -m gpu 2>&1 | tail -5
echo "--- sweep e2m1"
SWEEP_DTYPE=e2m1 SWEEP_CG=2 SWEEP_KBW=0,74 SWEEP_LEAD=0,2 timeout 300 python tools/sweep_gram.py 2>&1 | tail -8
echo "--- sweep i8 reference"
SWEEP_CG=2 SWEEP_KBW=0 SWEEP_LEAD=0 timeout 300 python tools/sweep_gram.py 2>&1 | tail -2
