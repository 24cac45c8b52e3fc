#!/bin/bash
# Retry gpurun while the pod answers "busy" (exit 3: nothing charged).  usage: tools/gpurun_retry.sh <timeout_s> '<command>' [--gpus N]
t=$1; cmd=$2; shift 2
for i in $(seq 1 60); do
  /usr/local/graft/bin/gpurun "$@" --timeout "$t" -- "$cmd"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 45
done
exit 3
