#!/bin/bash
# usage: tools/gpu_retry.sh <logfile> <gpurun args...>   -- retries while the pod answers "transient" (busy)
log=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  if ! grep -q "status=transient" "$log"; then exit 0; fi
  sleep 45
done
exit 3
