#!/bin/bash
# gpurun with retries while the pod answers "transient" (no slot free; nothing charged).  usage: tools/gpurun_retry.sh [gpurun args] -- 'cmd'
for i in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  if echo "$out" | grep -q "status=transient"; then sleep 45; continue; fi
  echo "$out"; exit 0
done
echo "$out"; echo "gave up after 30 transient answers"; exit 3
