#!/bin/bash
# gpurun with retries while the pod has no free slot (exit code 3 / "transient": nothing is charged).  usage: gpurun_retry.sh <log> <gpurun args...>
log=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  if ! grep -q "status=transient\|no box or slot" "$log"; then exit 0; fi
  sleep 90
done
exit 3
