#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu_gpu.py -x -q -m gpu 2>&1 | tail -3
for mode in nccl fused scatter:push scatter:pull scatter:copy; do
  red=${mode%%:*}; how=${mode##*:}
  VPCA_GATHER=$how timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 20 --warmup 3 --reduce $red --e2e-steps 0 --no-cpu-baseline --no-alt --no-eig-check \
    > gpurun_out/bench_2gpu_$mode.json 2> gpurun_out/bench_2gpu_$mode.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_2gpu_$mode.json').read().strip().splitlines()[-1])
print('$mode', '%.3e'%d['value'], round(d['ms_per_step'],3), 'close_ms', d.get('fused_close_ms'), d['checks'])
PY
done
