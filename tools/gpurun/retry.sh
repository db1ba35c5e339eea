#!/bin/bash
# usage: tools/gpurun/retry.sh <timeout-seconds> <out-file> <command...>   (retries while gpurun answers 3 = no box free)
T=$1; OUT=$2; shift 2
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun --timeout "$T" -- "$@" > "$OUT" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 150
done
exit 3
