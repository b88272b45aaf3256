#!/bin/bash
# usage: gpu_retry.sh <logfile> <timeout> <command...>   -- retries gpurun while the pod answers "busy" (exit 3)
log=$1; shift; to=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if ! grep -q "status=transient" $log; then exit $rc; fi
  sleep 60
done
exit 3
