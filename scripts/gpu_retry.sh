#!/bin/bash
# usage: scripts/gpu_retry.sh <timeout-seconds> '<command>'   -- retries gpurun while the pod answers busy / transient
T=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q '"status": *"transient"' gpurun_out/.last_call.json 2>/dev/null; then exit $rc; fi
  sleep 90
done
exit 3
