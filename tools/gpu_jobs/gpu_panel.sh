#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gram_gpu.py tests/test_driver_gpu.py -x -q -m gpu 2>&1 | tail -6
echo "--- int8 1M panels sweep"
SWEEP_PANEL=0,4096,8192,16384,32768 SWEEP_CG=2 SWEEP_KBW=0 SWEEP_REPS=8 python tools/sweep_gram.py 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    try: d=json.loads(l)
    except Exception: print(l.strip()); continue
    print({k:d[k] for k in ('v','panel','ms_med','ms_min','tops_syrk')}, d['prof']['mma_done_us_min_med_max'], d['ms_seq'][-4:])
"
echo "--- adaptive off int8 panel 8192"
VPCA_ADAPTIVE=0 SWEEP_PANEL=8192 SWEEP_CG=2 SWEEP_KBW=0 SWEEP_REPS=4 python tools/sweep_gram.py 2>&1 | tail -1 | cut -c1-330
echo "--- int8 5M panels"
SWEEP_V=5000000 SWEEP_PANEL=8192,16384 SWEEP_CG=2 SWEEP_KBW=0 SWEEP_REPS=5 python tools/sweep_gram.py 2>&1 | tail -2 | cut -c1-330
echo "--- bf16 1M panels"
SWEEP_DTYPE=bf16 SWEEP_PANEL=4096,8192 SWEEP_CG=2 SWEEP_KBW=0 SWEEP_REPS=5 python tools/sweep_gram.py 2>&1 | tail -2 | cut -c1-330
echo "--- e2m1 1M panels"
SWEEP_DTYPE=e2m1 SWEEP_PANEL=0,8192,16384 SWEEP_CG=2 SWEEP_KBW=0 SWEEP_REPS=5 python tools/sweep_gram.py 2>&1 | tail -3 | cut -c1-330
