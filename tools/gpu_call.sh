#!/bin/bash
# One gpurun call of round 2 (kept for reproducibility): the 4-GPU bench line (completes the 1 / 2 / 4 / 8 table).
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 20 --warmup 5 --no-cpu-baseline --no-also > gpurun_out/r2_bench_n4.json 2> gpurun_out/r2_bench_n4.err
grep -o '"value": [0-9.]*, "unit": "images/s", "n_gpus": 4, "steps": 20, "warmup": 5, "ms_per_step": [0-9.]*' gpurun_out/r2_bench_n4.json | head -1; tail -2 gpurun_out/r2_bench_n4.err
