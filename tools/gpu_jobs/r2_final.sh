#!/bin/bash
# round 2, final tree: the full default bench line on one GPU, then the ncu evidence behind it (captures tied to the final kernel source):
#   (a) launch list of a short bench run (shares of the step per kernel),
#   (b) ncu --set full of one warmed-up Gram launch per dtype (DRAM bytes, tensor pipe, L2),
#   (c) ncu --set full of the persistent Lanczos kernel + launch list of a whole vpca_compute_pca.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pca_gpu.py tests/test_parity_fullsize_gpu.py -q -m gpu --maxfail=5 --tb=short 2>&1 | tail -4 | cut -c1-300
VPCA_LZ_PROF=1 EIG_N=2504 EIG_MODES=auto EIG_REPS=7 timeout 300 python tools/eig_bench.py 2>&1 | tail -2
EIG_N=1092,4096,10000 EIG_MODES=auto EIG_REPS=5 timeout 300 python tools/eig_bench.py 2>&1 | tail -3
timeout 900 python bench.py > gpurun_out/r2_bench_final_1gpu.json 2> gpurun_out/r2_bench_final_1gpu.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/r2_bench_final_1gpu.json; tail -3 gpurun_out/r2_bench_final_1gpu.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_final_reference.json 2> gpurun_out/r2_bench_final_reference.err
echo "reference rc=$?"; tail -c 600 gpurun_out/r2_bench_final_reference.json
# (a)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches_final.csv \
   python bench.py --steps 4 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-legs > gpurun_out/r2_ncu_launch.log 2>&1
tail -2 gpurun_out/r2_ncu_launch.log | cut -c1-300
# (b) the split adapts over the first launches: skip them
for dt in i8 e2m1; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gram_kernel -s 24 -c 1 -f -o gpurun_out/r2_gram_final_$dt \
     python bench.py --dtype $dt --steps 30 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-eig-check --no-alt --no-legs > gpurun_out/r2_ncu_gram_$dt.log 2>&1
  tail -1 gpurun_out/r2_ncu_gram_$dt.log | cut -c1-200
done
# (c)
EIG_N=2504 EIG_MODES=auto EIG_REPS=1 timeout 500 ncu --set full --clock-control none --import-source on \
   -k regex:lz_persist_kernel -s 2 -c 1 -f -o gpurun_out/r2_lz_persist_final python tools/eig_bench.py > gpurun_out/r2_ncu_lz.log 2>&1
tail -2 gpurun_out/r2_ncu_lz.log | cut -c1-300
EIG_N=2504 EIG_MODES=auto EIG_REPS=1 timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none \
   -k regex:'lz_|bisect|invit|center|rowsum|matrix_mean' -c 200 --csv --log-file gpurun_out/r2_launches_eig_final.csv python tools/eig_bench.py > gpurun_out/r2_ncu_lz2.log 2>&1
for f in gpurun_out/r2_gram_final_i8 gpurun_out/r2_gram_final_e2m1 gpurun_out/r2_lz_persist_final; do
  [ -f $f.ncu-rep ] || continue
  ncu -i $f.ncu-rep --page raw --csv > ${f}_raw.csv 2>/dev/null
  sz=$(stat -c %s $f.ncu-rep); echo "$f.ncu-rep $sz bytes"
  if [ $sz -gt 16000000 ]; then rm -f $f.ncu-rep; fi     # gpurun_out/ is capped at 64 MiB: keep the raw page, drop big reports
done
ls -la gpurun_out | tail -20
