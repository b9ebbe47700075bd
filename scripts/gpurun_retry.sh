#!/bin/bash
# usage: gpurun_retry.sh <log> <timeout> <script>   -- retries while the pod reports "transient" (nothing charged)
log=$1; to=$2; shift 2
for i in 1 2 3 4 5 6 7 8; do
  /usr/local/graft/bin/gpurun --timeout $to -- "$@" > $log 2>&1
  if ! grep -q "status=transient" $log; then break; fi
  sleep 60
done
