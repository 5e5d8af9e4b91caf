#!/bin/bash
# round 2, last call: the whole GPU suite + smoke on the final tree, eigensolver timings with the packed one-reduction pass
mkdir -p gpurun_out
VPCA_LZ_PROF=1 EIG_N=2504 EIG_MODES=auto EIG_REPS=7 timeout 300 python tools/eig_bench.py 2>&1 | tail -2
EIG_N=1092,4096,10000 EIG_MODES=auto EIG_REPS=5 timeout 300 python tools/eig_bench.py 2>&1 | tail -3
timeout 1700 python -m pytest tests -q -m gpu --maxfail=8 --tb=short > gpurun_out/r2_fullsuite2.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/r2_fullsuite2.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-legs --e2e-steps 2 > gpurun_out/r2_bench_last.json 2> gpurun_out/r2_bench_last.err
echo "bench rc=$?"; python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench_last.json").read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], "kernel_ms", d["roofline"]["kernel_ms"], "traffic", d["roofline"]["traffic"], "eig_ms", d["eig_ms"], "checks", d["checks"], "e2e", d["e2e"]["ms_per_step"])
PY
