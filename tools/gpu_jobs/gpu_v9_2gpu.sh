#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu_gpu.py -x -q -m gpu 2>&1 | tail -5
for mode in nccl fused scatter; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus 2 --steps 20 --warmup 3 --reduce $mode --e2e-steps 0 --no-cpu-baseline --no-alt --no-eig-check \
    > gpurun_out/bench_2gpu_$mode.json 2> gpurun_out/bench_2gpu_$mode.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/bench_2gpu_$mode.json').read().strip().splitlines()[-1])
print('$mode', d['value'], d['ms_per_step'], d['checks'])
PY
done
