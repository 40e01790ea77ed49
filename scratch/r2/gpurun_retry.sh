#!/bin/bash
# usage: gpurun_retry.sh [gpurun args...] -- retries while the pod answers "transient" (busy slots)
for i in $(seq 1 12); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 150; continue; fi
  echo "$out" | tail -40
  exit 0
done
echo "gave up: pod busy"
