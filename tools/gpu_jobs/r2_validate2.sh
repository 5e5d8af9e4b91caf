#!/bin/bash
# round 2, call 2: after the TMEM-column fixes (mxf4 accumulators at 0 / 240, square tiling back to 512 columns), the
# repaired speed-weighted split, and balanced strip widths: full GPU test suite, per-launch adaptation traces, A/B bench
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 1500 python -m pytest tests -q -m gpu --maxfail=6 2>&1 | tail -40
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for ex in 1 0; do
  echo "=== sweep int8 exact_cover=$ex (ms_seq = per-launch ms: the split adapts from launch to launch) ==="
  VPCA_EXACT_COVER=$ex SWEEP_CG=2 SWEEP_KBW=0 SWEEP_REPS=14 timeout 300 python tools/sweep_gram.py 2>&1 | tail -3
done
echo "=== sweep e2m1 (mxf4) ==="
SWEEP_DTYPE=e2m1 SWEEP_CG=2 SWEEP_KBW=0 SWEEP_REPS=14 timeout 300 python tools/sweep_gram.py 2>&1 | tail -3
echo "=== sweep int8 exact, no adaptation ==="
VPCA_ADAPTIVE=0 SWEEP_CG=2 SWEEP_KBW=0 SWEEP_REPS=6 timeout 300 python tools/sweep_gram.py 2>&1 | tail -2
for ex in 0 1; do
  VPCA_EXACT_COVER=$ex timeout 300 python bench.py --steps 20 --warmup 5 --e2e-steps 0 --no-cpu-baseline --no-eig-check \
     > gpurun_out/r2_bench_ab2_exact$ex.json 2> gpurun_out/r2_bench_ab2_exact$ex.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_ab2_exact$ex.json").read().strip().splitlines()[-1])
    print("exact_cover=$ex", "ms/step", round(d["ms_per_step"], 4), "kernel_ms", round(d["roofline"]["kernel_ms"], 4), "checks", d["checks"],
          "mxf4", round(d.get("packed_e2m1", {}).get("kernel_ms", 0), 4), d.get("packed_e2m1", {}).get("gram_bit_identical_to_int8_path"), "clocks", d["clocks"])
except Exception as exc:
    print("exact_cover=$ex bench failed:", exc); print(open("gpurun_out/r2_bench_ab2_exact$ex.err").read()[-1500:])
PY
done
