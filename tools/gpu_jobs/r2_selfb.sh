#!/bin/bash
# round 2, last kernel change: diagonal 256 x 256 tiles read B through the A tile (no B load).  Parity of the Gram paths,
# A/B against VPCA_SELF_B=0, and -- because the kernel code changed -- fresh ncu captures for profiles/r2_gram_traffic.json
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gram_gpu.py -q -m gpu -x --tb=short 2>&1 | tail -4 | cut -c1-300
for sb in 1 0 1 0; do
  echo "VPCA_SELF_B=$sb"
  VPCA_SELF_B=$sb SWEEP_PANEL=8192 SWEEP_CG=2 SWEEP_KBW=0 SWEEP_REPS=12 timeout 200 python tools/sweep_gram.py 2>&1 | tail -1 | python -c "import sys, json; d = json.loads(sys.stdin.readline()); print(d['ms_seq'][-8:], 'med', d['ms_med'], 'min', d['ms_min'], d['prof'], d['checksum'])"
done
for dt in i8 e2m1; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:gram_kernel -s 24 -c 1 -f -o gpurun_out/r2_gram_selfb_$dt \
     python bench.py --dtype $dt --steps 30 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-eig-check --no-alt --no-legs > gpurun_out/r2_ncu_gram_selfb_$dt.log 2>&1
  ncu -i gpurun_out/r2_gram_selfb_$dt.ncu-rep --page raw --csv > gpurun_out/r2_gram_selfb_${dt}_raw.csv 2>/dev/null
  rm -f gpurun_out/r2_gram_selfb_$dt.ncu-rep
done
timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-legs --e2e-steps 0 --no-eig-check > gpurun_out/r2_bench_selfb.json 2>/dev/null
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_selfb.json").read().strip().splitlines()[-1])
print("bench: ms/step", d["ms_per_step"], "kernel_ms", d["roofline"]["kernel_ms"], "checks", d["checks"], "mxf4", d.get("packed_e2m1", {}).get("kernel_ms"), d.get("packed_e2m1", {}).get("gram_bit_identical_to_int8_path"))
PY
