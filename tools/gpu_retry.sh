#!/bin/bash
# usage: tools/gpu_retry.sh <logfile> <timeout_s> <command...>   — retries gpurun while the pod answers "busy" (exit 3)
log=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout "$to" -- "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ] && ! grep -q "status=transient" "$log"; then exit $rc; fi
  sleep 45
done
exit 3
