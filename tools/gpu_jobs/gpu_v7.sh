#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_pca_gpu.py tests/test_driver_gpu.py -x -q -m gpu 2>&1 | tail -12
EIG_N=1092,2504,4096,8192 timeout 600 python tools/eig_bench.py 2>&1 | tail -10 | tee gpurun_out/eig_bench_v7.jsonl
EIG_N=2504 EIG_K=6 timeout 300 python tools/eig_bench.py 2>&1 | tail -2 | tee -a gpurun_out/eig_bench_v7.jsonl
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-alt > gpurun_out/bench_v7.json 2> gpurun_out/bench_v7.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_v7.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','eig_ms','eig')}); e=d['e2e']; print({k:e.get(k) for k in ('value','ms_per_step','h2d_bytes_per_step','error')}); print('u16',e.get('with_uint16_indices')); print('bits',e.get('with_bitmap_rows')); print(d['checks'])
PY
tail -3 gpurun_out/bench_v7.err
