#!/bin/bash
# One gpurun call of round 2 (kept for reproducibility): step tests after moving the update-independent loss values ahead of the D all-reduce wait.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_steps_gpu.py tests/test_headline_gpu.py -q -x > gpurun_out/r2_step_tests.log 2>&1; echo rc=$?; tail -3 gpurun_out/r2_step_tests.log
