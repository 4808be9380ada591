#!/bin/bash
# One gpurun call of round 2 (kept for reproducibility): full GPU suite at HEAD (test order changed by the global_d tests).
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu.log
