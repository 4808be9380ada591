#!/bin/bash
# One gpurun call of round 2 (kept for reproducibility): ncu --set full of the 1-channel-side kernels inside a train step.
mkdir -p gpurun_out
python tools/one_step.py 2 64 256 wgan > gpurun_out/one_step.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:c1_ -s 11 -c 11 -f -o gpurun_out/r2_prof_c1 python tools/one_step.py 2 64 256 wgan > gpurun_out/ncu_c1.log 2>&1
tail -2 gpurun_out/ncu_c1.log
