#!/bin/bash
# One gpurun call of round 2 (kept for reproducibility): max-pool fused into the 3x3 convolution epilogue — tests, per-layer table, config-4 bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_vgg_gpu.py -q -x > gpurun_out/r2_vgg_tests.log 2>&1; echo rc=$?; tail -8 gpurun_out/r2_vgg_tests.log
timeout 300 python tools/vgg_table.py > gpurun_out/r2_vgg_table.txt 2>&1; tail -34 gpurun_out/r2_vgg_table.txt
timeout 600 python bench.py --recipe wgan_perceptual_style --steps 20 --warmup 5 > gpurun_out/r2_bench_config4_n1.json 2> gpurun_out/r2_bench_config4_n1.err; cut -c1-300 gpurun_out/r2_bench_config4_n1.json; tail -3 gpurun_out/r2_bench_config4_n1.err
