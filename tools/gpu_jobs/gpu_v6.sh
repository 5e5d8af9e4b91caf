#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gram_gpu.py tests/test_driver_gpu.py tests/test_pca_gpu.py -x -q -m gpu -k "bitmap or checkpoint or top2 or end_to_end or main_sequence" 2>&1 | tail -6
python tools/eig_bench.py 2>&1 | tail -3
python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-alt > gpurun_out/bench_v6.json 2> gpurun_out/bench_v6.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_v6.json').read().strip().splitlines()[-1])
print({k:d[k] for k in ('value','ms_per_step','eig_ms')}); e=d['e2e']; print({k:e[k] for k in ('value','ms_per_step','h2d_bytes_per_step')}); print('u16',e['with_uint16_indices']); print('bits',e['with_bitmap_rows'])
PY
tail -3 gpurun_out/bench_v6.err
