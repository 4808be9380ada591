#!/bin/bash
# One gpurun call of round 2 (kept for reproducibility): final 8-GPU bench lines (headline, config 4) with the round's last kernels.
mkdir -p gpurun_out
run() {  # name, extra flags
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 30 --warmup 5 \
    --no-cpu-baseline --no-also $2 > gpurun_out/r2_bench_$1.json 2> gpurun_out/r2_bench_$1.err
  grep -o '"value": [0-9.]*, "unit": "images/s", "n_gpus": 8, "steps": 30, "warmup": 5, "ms_per_step": [0-9.]*' gpurun_out/r2_bench_$1.json | head -1 | sed "s/^/$1 /"
}
run n8 ""
run config4_n8 "--recipe wgan_perceptual_style"
