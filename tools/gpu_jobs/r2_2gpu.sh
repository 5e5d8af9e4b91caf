#!/bin/bash
# round 2, 2 GPUs: the multi-GPU tests for real (NCCL + fused peer reduce across two devices, pool across two devices),
# the eigensolver tests after the 8-step deflated re-run, and the bench at N = 2 with all legs
mkdir -p gpurun_out
nvidia-smi -L
python - <<'PY'
from spark_examples_b200 import native
print("clusters of the Gram kernel the device holds at once (1 CTA per SM):", {c: native.maxClusters(0, c) for c in (1, 2, 4, 8, 16)})
PY
timeout 900 python -m pytest tests/test_multigpu_gpu.py tests/test_pool_gpu.py tests/test_pca_gpu.py -q -m gpu --maxfail=5 --tb=short 2>&1 | tail -15 | cut -c1-300
echo "=== eig: S rows resident in shared memory (default) vs streamed from L2 (VPCA_LZ_SROWS=0) ==="
VPCA_LZ_PROF=1 EIG_N=2504 EIG_MODES=auto EIG_REPS=7 timeout 300 python tools/eig_bench.py 2>&1 | tail -2
VPCA_LZ_SROWS=0 VPCA_LZ_PROF=1 EIG_N=2504 EIG_MODES=auto EIG_REPS=7 timeout 300 python tools/eig_bench.py 2>&1 | tail -2
EIG_N=1092,4096,10000 EIG_MODES=auto EIG_REPS=5 timeout 300 python tools/eig_bench.py 2>&1 | tail -3
echo "=== Gram A/B (panels of 8192, int8): 64-bit reds and rebalance gain ==="
for cfg in "1 0.7" "0 0.7" "1 0.5" "1 1.0"; do set -- $cfg
  echo "VPCA_RED64=$1 VPCA_REBALANCE_GAIN=$2"
  VPCA_RED64=$1 VPCA_REBALANCE_GAIN=$2 SWEEP_PANEL=8192 SWEEP_CG=2 SWEEP_KBW=0 SWEEP_REPS=16 timeout 300 python tools/sweep_gram.py 2>&1 | tail -1 | python -c "import sys, json; d = json.loads(sys.stdin.readline()); print(d['ms_seq'], 'med', d['ms_med'], 'min', d['ms_min'], d['prof'])"
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 \
   > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
echo "bench n=2 rc=$?"; tail -c 3000 gpurun_out/r2_bench_n2.json; tail -5 gpurun_out/r2_bench_n2.err | cut -c1-300
