#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/large_n_shard.py --samples 100000 --variants 62500 > gpurun_out/c4_shard.json 2> gpurun_out/c4_shard.err
echo rc=$?; cat gpurun_out/c4_shard.json; tail -5 gpurun_out/c4_shard.err
