#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <gpurun args...>   (retries while the pod answers "busy")
log=$1; shift
for i in 1 2 3 4 5 6 7 8 9 10 11 12; do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  if grep -q "status=transient\|exit code 3\|status=busy" "$log"; then sleep 100; continue; fi
  break
done
