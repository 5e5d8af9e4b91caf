#!/bin/bash
# usage: tools/gpurun_retry.sh <log> <timeout_s> [--gpus N] -- <command>
# Retries while the pod answers "busy / draining" (nothing is charged for those), up to ~40 min.
log=$1; shift; tmo=$1; shift
extra=()
while [ "$1" != "--" ]; do extra+=("$1"); shift; done
shift
for attempt in $(seq 1 14); do
  /usr/local/graft/bin/gpurun --timeout "$tmo" "${extra[@]}" -- "$@" > "$log" 2>&1
  rc=$?
  if grep -q "status=transient" "$log" || [ $rc -eq 3 ]; then sleep 170; continue; fi
  exit $rc
done
exit 3
