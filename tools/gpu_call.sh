#!/bin/bash
# One gpurun call of round 2 (kept for reproducibility): VGG perceptual / style loss kernels — tests and per-layer timing.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_vgg_gpu.py -q > gpurun_out/r2_vgg_tests.log 2>&1; echo rc=$?; tail -15 gpurun_out/r2_vgg_tests.log
timeout 300 python tools/vgg_table.py > gpurun_out/r2_vgg_table.txt 2>&1; cat gpurun_out/r2_vgg_table.txt | tail -50
