#!/bin/bash
# usage: gpurun_retry.sh <log> <timeout> [--gpus N] <command>   -- retries while the pod reports "transient" / busy (nothing charged)
log=$1; to=$2; shift 2
extra=""
if [ "$1" = "--gpus" ]; then extra="--gpus $2"; shift 2; fi
for i in 1 2 3 4 5 6 7 8 9 10; do
  /usr/local/graft/bin/gpurun $extra --timeout $to -- "$@" > $log 2>&1
  rc=$?
  if ! grep -q "status=transient" $log && [ $rc -ne 3 ]; then break; fi
  sleep 60
done
