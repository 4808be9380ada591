#!/bin/bash
# One gpurun call of round 2 (kept for reproducibility): ncu --set full of the 64 -> 128 stride-2 convolution with the statistics epilogue (G d2 / D c2).
mkdir -p gpurun_out
timeout 600 ncu --set full --import-source on --clock-control none -k regex:igemm -s 3 -c 1 -o gpurun_out/r2_conv_d2 -f python tools/one_layer.py conv 64 128 64 128 5 1 > gpurun_out/ncu_d2.log 2>&1; tail -2 gpurun_out/ncu_d2.log
