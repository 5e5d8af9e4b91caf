#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multigpu_gpu.py -x -q -m gpu 2>&1 | tail -15
for mode in nccl fused; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
   bench.py --gpus $N --steps 20 --warmup 3 --reduce $mode --e2e-steps 2 --no-alt > gpurun_out/bench_n${N}_$mode.json 2> gpurun_out/bench_n${N}_$mode.err
python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_n${N}_$mode.json').read())
    print('$mode', {k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['config']['reduce'], d['checks'], d.get('e2e',{}).get('value'))
except Exception as e:
    print('$mode failed', e)
PY
tail -4 gpurun_out/bench_n${N}_$mode.err
done
