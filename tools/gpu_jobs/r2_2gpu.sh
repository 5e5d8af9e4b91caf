#!/bin/bash
# round 2, 2 GPUs: the multi-GPU tests for real (NCCL + fused peer reduce across two devices, pool across two devices),
# the eigensolver tests after the 8-step deflated re-run, and the bench at N = 2 with all legs
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_multigpu_gpu.py tests/test_pool_gpu.py tests/test_pca_gpu.py -q -m gpu --maxfail=5 --tb=short 2>&1 | tail -15 | cut -c1-300
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 \
   > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
echo "bench n=2 rc=$?"; tail -c 3000 gpurun_out/r2_bench_n2.json; tail -5 gpurun_out/r2_bench_n2.err | cut -c1-300
