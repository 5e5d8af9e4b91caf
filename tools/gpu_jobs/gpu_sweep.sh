#!/bin/bash
mkdir -p gpurun_out
SWEEP_CG=2 python tools/sweep_gram.py > gpurun_out/sweep2.jsonl 2> gpurun_out/sweep2.err
cat gpurun_out/sweep2.jsonl; tail -5 gpurun_out/sweep2.err
