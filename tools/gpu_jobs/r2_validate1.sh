#!/bin/bash
# round 2, call 1: the refactored boundary (lanes, pool, local peers, band-only Grams, C harnesses), the exact-block-cover
# Gram schedule, and an A/B bench of the two tilings
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -25
if [ ${PIPESTATUS[0]} -ne 0 ]; then
  echo "=== retry of the Gram tests with the square tiling (VPCA_EXACT_COVER=0) ==="
  VPCA_EXACT_COVER=0 timeout 900 python -m pytest tests/test_gram_gpu.py tests/test_pool_gpu.py -x -q -m gpu 2>&1 | tail -15
fi
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
for ex in 0 1; do
  VPCA_EXACT_COVER=$ex timeout 300 python bench.py --steps 20 --warmup 3 --e2e-steps 0 --no-cpu-baseline --no-eig-check \
     > gpurun_out/r2_bench_ab_exact$ex.json 2> gpurun_out/r2_bench_ab_exact$ex.err
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/r2_bench_ab_exact$ex.json").read().strip().splitlines()[-1])
    print("exact_cover=$ex", "ms/step", round(d["ms_per_step"], 4), "kernel_ms", round(d["roofline"]["kernel_ms"], 4), "checks", d["checks"],
          "mxf4", round(d.get("packed_e2m1", {}).get("kernel_ms", 0), 4), "clocks", d["clocks"])
except Exception as exc:
    print("exact_cover=$ex bench failed:", exc); print(open("gpurun_out/r2_bench_ab_exact$ex.err").read()[-1500:])
PY
done
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_v1.json 2> gpurun_out/r2_bench_v1.err
tail -c 2500 gpurun_out/r2_bench_v1.json; tail -5 gpurun_out/r2_bench_v1.err
