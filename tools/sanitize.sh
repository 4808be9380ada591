#!/bin/bash
# compute-sanitizer over one small train step of each recipe family (128x128, batch 4: every kernel of the hot path runs, CTA pairs
# and the window operand staging included).  ONE tool per gpurun call (B200_PROFILING.md):   tools/sanitize.sh memcheck|racecheck|synccheck|initcheck
TOOL=${1:-memcheck}
mkdir -p gpurun_out
compute-sanitizer --tool "$TOOL" --print-limit 20 python tools/one_step.py 1 4 128 wgan > gpurun_out/sanitizer_${TOOL}_wgan.log 2>&1
echo "[sanitize] $TOOL wgan rc=$?"; tail -4 gpurun_out/sanitizer_${TOOL}_wgan.log
