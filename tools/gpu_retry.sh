#!/bin/bash
# usage: tools/gpu_retry.sh TIMEOUT 'command'   -- re-submits while the pod answers busy (exit 3 / "transient"), up to 12 times
t=$1; shift
for i in $(seq 1 12); do
  out=$(/usr/local/graft/bin/gpurun --timeout "$t" -- "$@" 2>&1)
  echo "$out" | tail -25
  if echo "$out" | grep -q "status=transient\|no box or slot"; then sleep 150; continue; fi
  break
done
