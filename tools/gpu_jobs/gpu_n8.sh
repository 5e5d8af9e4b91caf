#!/bin/bash
N=8
export VPCA_EIG_TWO_KERNELS=1   # validated eigensolve path for this (expensive) run
mkdir -p gpurun_out
nvidia-smi -L | head -8
run() { # name, extra args
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
     bench.py --gpus $N --steps 20 --warmup 3 $2 > gpurun_out/bench_n8_$1.json 2> gpurun_out/bench_n8_$1.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/bench_n8_$1.json').read().strip().splitlines()[-1])
    print('$1', {k:d[k] for k in ('value','ms_per_step','n_gpus')}, d['config']['reduce'][:30], d['checks'], 'roof', round(d['roofline']['frac'],3), 'kernel_ms', round(d['roofline']['kernel_ms'],3), d['clocks'], 'e2e', d.get('e2e',{}).get('value'))
except Exception as e:
    print('$1 failed', e)
PY
  grep -v "OMP_NUM_THREADS\|^\*\*\*\|^$" gpurun_out/bench_n8_$1.err | tail -4
}
run nccl "--e2e-steps 2"
run fused "--reduce fused --e2e-steps 0 --no-alt"
run c3_nccl "--variants-per-gpu 5000000 --e2e-steps 0 --no-alt"
run c3_e2m1 "--variants-per-gpu 5000000 --e2e-steps 0 --dtype e2m1"
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus $N --steps 2 --warmup 1 --cpu-seconds 4 2>/dev/null | tail -1 | cut -c1-300
