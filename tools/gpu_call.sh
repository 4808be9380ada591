#!/bin/bash
# One gpurun call of round 2 (kept for reproducibility): ncu --set full of the first 3x3 tensor-core layer (conv1_2, 64 -> 64 at 256x256, 64 images).
mkdir -p gpurun_out
timeout 300 python tools/vgg_table.py --reps 1 > gpurun_out/vgg_pre.txt 2>&1; tail -1 gpurun_out/vgg_pre.txt
timeout 600 ncu --set full --import-source on --clock-control none -k regex:igemm -c 1 -o gpurun_out/r2_conv3_c64 -f python tools/vgg_table.py --reps 1 > gpurun_out/ncu_conv3.log 2>&1; tail -2 gpurun_out/ncu_conv3.log
ls -la gpurun_out/*.ncu-rep
