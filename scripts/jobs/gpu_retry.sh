#!/bin/bash
# usage: gpu_retry.sh <log> <timeout> <command...>   -- retries while the pod answers "busy" (exit code 3), nothing is charged for those
log=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
