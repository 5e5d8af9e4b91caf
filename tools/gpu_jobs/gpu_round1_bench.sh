#!/bin/bash
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 3 > gpurun_out/bench_first.json 2> gpurun_out/bench_first.err
tail -c 4000 gpurun_out/bench_first.json; tail -5 gpurun_out/bench_first.err
python tools/sweep_gram.py > gpurun_out/sweep_first.jsonl 2> gpurun_out/sweep_first.err
cat gpurun_out/sweep_first.jsonl; tail -5 gpurun_out/sweep_first.err
