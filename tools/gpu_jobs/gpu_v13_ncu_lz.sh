#!/bin/bash
mkdir -p gpurun_out
EIG_N=2504 EIG_MODES=auto EIG_REPS=1 timeout 500 ncu --set full --clock-control none --import-source on \
   -k regex:lz_matvec_kernel -s 40 -c 2 -o gpurun_out/lz_matvec -f python tools/eig_bench.py > gpurun_out/ncu_lz.log 2>&1
tail -3 gpurun_out/ncu_lz.log
EIG_N=2504 EIG_MODES=auto EIG_REPS=1 timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none \
   -k regex:'lz_|bisect|invit|center|rowsum|matrix_mean' -c 500 --csv --log-file gpurun_out/launches_eig.csv python tools/eig_bench.py > gpurun_out/ncu_lz2.log 2>&1
tail -2 gpurun_out/ncu_lz2.log
ls -la gpurun_out/lz_matvec.ncu-rep
