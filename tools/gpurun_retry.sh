#!/bin/bash
# usage: tools/gpurun_retry.sh <logfile> <gpurun args...>   -- retries while the pod answers busy / transient (exit 3), nothing is charged for those
log=$1; shift
for i in $(seq 1 40); do
  /usr/local/graft/bin/gpurun "$@" > "$log" 2>&1
  rc=$?
  if [ $rc -ne 3 ]; then echo "gpurun rc=$rc" >> "$log"; exit $rc; fi
  sleep 45
done
echo "gpurun: gave up after 40 busy answers" >> "$log"
