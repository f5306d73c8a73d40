#!/bin/bash
# Retry wrapper around gpurun for background use: retries while the pod answers "busy" (exit 3), up to 12 times.
# usage: tools/gpu_call.sh <timeout-seconds> '<command>' [--gpus N]
T=$1; CMD=$2; shift 2
for i in $(seq 1 12); do
  /usr/local/graft/bin/gpurun "$@" --timeout "$T" -- "$CMD"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 90
done
exit 3
