#!/bin/bash
# Retry wrapper around gpurun: exit code 3 (no box / slot free) is retried every 2 minutes.
#   tools/gpu.sh [--gpus N] <timeout_s> '<command>'
GP=""
if [ "$1" == "--gpus" ]; then GP="--gpus $2"; shift 2; fi
T=$1; shift
for i in $(seq 1 30); do
  /usr/local/graft/bin/gpurun $GP --timeout "$T" -- "$@"
  rc=$?
  if [ $rc -ne 3 ]; then exit $rc; fi
  sleep 120
done
exit 3
