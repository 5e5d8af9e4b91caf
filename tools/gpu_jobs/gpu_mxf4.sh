#!/bin/bash
mkdir -p gpurun_out
export VPCA_E2M1_MXF4=1
for t in "test_gram_e2m1_dense_and_calls[2]" "test_gram_e2m1_dense_and_calls[1]" "test_synth_and_gram_panel_layout[e2m1]" test_gram_e2m1_resident_matches_int8; do
  timeout 120 python -m pytest "tests/test_gram_gpu.py::$t" -x -q -m gpu > gpurun_out/mxf4_$$.log 2>&1; rc=$?
  echo "$t rc=$rc"; if [ $rc -ne 0 ]; then tail -30 gpurun_out/mxf4_$$.log; fi
done
echo "--- sweep e2m1 via mxf4"
SWEEP_DTYPE=e2m1 SWEEP_PANEL=8192,16384 SWEEP_CG=2 SWEEP_KBW=0 SWEEP_REPS=6 timeout 300 python tools/sweep_gram.py 2>&1 | tail -3 | cut -c1-420
unset VPCA_E2M1_MXF4
echo "--- sweep e2m1 via f8f6f4 and int8 (n_eff trimming on)"
SWEEP_DTYPE=e2m1 SWEEP_PANEL=8192 SWEEP_CG=2 SWEEP_KBW=0 SWEEP_REPS=6 timeout 300 python tools/sweep_gram.py 2>&1 | tail -1 | cut -c1-420
SWEEP_PANEL=8192 SWEEP_CG=2 SWEEP_KBW=0 SWEEP_REPS=6 timeout 300 python tools/sweep_gram.py 2>&1 | tail -1 | cut -c1-420
