#!/bin/bash
# round 2, 8 GPUs: BASELINE configs[3] in the owner-computes form (no peers, no traffic between the GPUs), next to the
# owner-flush form of r2_8gpu.sh on the same box
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python tools/c4_bands.py --mode owner-computes --reps 3 --check-rows 2 --out gpurun_out/r2_c4_8gpu_owner_computes.json 2> gpurun_out/r2_c4_computes.err | cut -c1-1500
echo "c4 owner-computes rc=${PIPESTATUS[0]}"; tail -3 gpurun_out/r2_c4_computes.err | cut -c1-300
