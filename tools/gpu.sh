#!/bin/bash
# Retry wrapper around gpurun: "no box / slot free" (exit code 3) and "transient" answers are retried every 2 minutes.
#   tools/gpu.sh [--gpus N] <timeout_s> '<command>'
GP=""
if [ "$1" == "--gpus" ]; then GP="--gpus $2"; shift 2; fi
T=$1; shift
for i in $(seq 1 30); do
  out=$(/usr/local/graft/bin/gpurun $GP --timeout "$T" -- "$@" 2>&1)
  rc=$?
  echo "$out"
  if [ $rc -ne 3 ] && ! echo "$out" | grep -q "status=transient"; then exit $rc; fi
  sleep 120
done
exit 3
