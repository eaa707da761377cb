#!/bin/bash
# gpurun with retries while the pod answers "transient / busy" (nothing is charged for those answers).
#   tools/gpurun_retry.sh [gpurun options] -- '<command>'
for attempt in $(seq 1 20); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  echo "$out"
  if echo "$out" | grep -q "status=transient\|rc=3\b\|no box or slot"; then
    echo "[retry] attempt $attempt answered busy; sleeping 90 s"
    sleep 90
    continue
  fi
  break
done
