#!/bin/bash
# usage: tools/gpu_retry.sh <timeout_s> <command...>   -- retries while gpurun answers "busy" (rc 3), up to 12 times
t=$1; shift
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun --timeout "$t" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
