#!/bin/bash
# One gpurun call of round 2 (kept for reproducibility): the global_d recipe (experiment3_global_D) — step tests, golden scalars, begin() surface.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_steps_gpu.py -q -x -k "global_d or global_local or begin_call" > gpurun_out/r2_global_d_tests.log 2>&1; echo rc=$?; tail -12 gpurun_out/r2_global_d_tests.log
