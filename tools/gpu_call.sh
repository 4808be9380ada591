#!/bin/bash
# One gpurun call of round 2 (kept for reproducibility): full GPU suite with the config-4 additions + the C4 parity table.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; echo rc=$?; tail -6 gpurun_out/r2_pytest_gpu.log
timeout 600 python -m tests.parity_report --cases C4 --out gpurun_out/parity_r2_c4.json > gpurun_out/parity_c4.log 2>&1; tail -3 gpurun_out/parity_c4.log
