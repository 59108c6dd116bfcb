#!/bin/bash
# gpurun with retries while the pod answers "busy / draining" (nothing is charged for those):
#   tools/gpurun_retry.sh <timeout seconds> '<command>'
for i in 1 2 3 4 5 6 7 8; do
  out=$(/usr/local/graft/bin/gpurun --timeout "$1" -- "$2" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 150; continue; fi
  echo "$out"; exit 0
done
echo "$out"
