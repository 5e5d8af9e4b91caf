#!/bin/bash
# round 2, 8 GPUs of one box: bench at N = 8 (headline 8 x 1 M + the c3 leg = BASELINE configs[2] 2504 x 40 M + the c5_bf16 leg
# = configs[4] point 8), configs[3] with band-only Grams from one process, and the same-process pool across 8 devices
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 3 \
   > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err
echo "bench n=8 rc=$?"; tail -c 2500 gpurun_out/r2_bench_n8.json; tail -4 gpurun_out/r2_bench_n8.err | cut -c1-300
timeout 900 python tools/c4_bands.py --out gpurun_out/r2_c4_8gpu.json 2> gpurun_out/r2_c4_8gpu.err | cut -c1-1500
echo "c4 rc=${PIPESTATUS[0]}"; tail -3 gpurun_out/r2_c4_8gpu.err | cut -c1-300
timeout 600 python -m pytest tests/test_pool_gpu.py tests/test_multigpu_gpu.py -q -m gpu --maxfail=3 --tb=short 2>&1 | tail -6 | cut -c1-300
