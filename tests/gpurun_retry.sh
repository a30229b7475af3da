#!/bin/bash
# usage: gpurun_retry.sh <logfile> <gpurun args...>   -- retries while gpurun answers "no box free" (exit 3)
log=$1; shift
for attempt in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
