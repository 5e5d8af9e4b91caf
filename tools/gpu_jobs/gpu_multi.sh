#!/bin/bash
# usage: tools/gpu_multi.sh N   -- bench at N GPUs via torchrun (the driver's launch line)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -c 2500 gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
   bench.py --impl reference --gpus $N --steps 2 --warmup 1 --cpu-seconds 3 > gpurun_out/bench_ref_n$N.json 2> gpurun_out/bench_ref_n$N.err
tail -c 600 gpurun_out/bench_ref_n$N.json; tail -3 gpurun_out/bench_ref_n$N.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 \
   -m spark_examples_b200.variants_pca --synthetic 2504,400000 --variants-per-partition 50000 > gpurun_out/driver_n$N.out 2> gpurun_out/driver_n$N.err
head -3 gpurun_out/driver_n$N.out; tail -4 gpurun_out/driver_n$N.out | cut -c1-300; tail -3 gpurun_out/driver_n$N.err
