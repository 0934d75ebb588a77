#!/bin/bash
# gpurun with retries while the pod answers "transient" (busy): tools/gpurun_retry.sh [gpurun args...] -- 'command'
for k in 1 2 3 4 5 6 7 8; do
  out=$(/usr/local/graft/bin/gpurun "$@" 2>&1)
  echo "$out" | tail -80
  if echo "$out" | grep -q "status=transient"; then sleep 90; continue; fi
  break
done
