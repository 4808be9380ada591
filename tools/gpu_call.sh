#!/bin/bash
# One gpurun call of round 2 (kept for reproducibility): ncu --set full of the 1-channel-side kernels inside one eager train step.
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"wgrad_c1_kernel|convt_c1_fwd|conv_c1_fwd" --launch-skip 11 -c 11 -o gpurun_out/r2_c1_kernels -f \
  python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-graph --no-also > gpurun_out/ncu_c1.log 2>&1; tail -2 gpurun_out/ncu_c1.log
ls -la gpurun_out/r2_c1_kernels.ncu-rep
