#!/bin/bash
mkdir -p gpurun_out
for mode in nccl scatter; do
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 8 --steps 30 --warmup 3 --reduce $mode --e2e-steps 0 --no-cpu-baseline --no-alt --no-eig-check \
    > gpurun_out/bench_8gpu_$mode.json 2> gpurun_out/bench_8gpu_$mode.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_8gpu_$mode.json').read().strip().splitlines()[-1])
    print('$mode', '%.3e'%d['value'], round(d['ms_per_step'],3), 'close_ms', d.get('fused_close_ms'), d['checks'], d['clocks'].get('reasons'))
except Exception as e:
    print('$mode failed', e)
PY
done
tail -3 gpurun_out/bench_8gpu_scatter.err
