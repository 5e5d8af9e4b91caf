#!/bin/bash
# Run every GPU test in its own process (a trapped kernel kills the CUDA context of its process only),
# each under a timeout, and collect a summary in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/probe_gpu.txt 2>&1
python -m pytest tests --collect-only -q -m gpu 2>/dev/null | grep "::" > gpurun_out/probe_ids.txt
: > gpurun_out/probe_summary.txt
while read -r id; do
  start=$(date +%s.%N)
  timeout 300 python -m pytest "$id" -x -q -m gpu > gpurun_out/probe_last.log 2>&1
  rc=$?
  end=$(date +%s.%N)
  printf "%s rc=%d %.1fs\n" "$id" "$rc" "$(echo "$end - $start" | bc)" >> gpurun_out/probe_summary.txt
  if [ $rc -ne 0 ]; then
    { echo "=== $id (rc=$rc)"; tail -60 gpurun_out/probe_last.log; } >> gpurun_out/probe_failures.txt
  fi
done < gpurun_out/probe_ids.txt
cat gpurun_out/probe_summary.txt
echo "---- failures (head)"
head -150 gpurun_out/probe_failures.txt 2>/dev/null
