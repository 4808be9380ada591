#!/bin/bash
# One gpurun call of round 2 (kept for reproducibility): bench.py with the median-of-three instrumented steps, 1 and 2 GPUs.
mkdir -p gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; cut -c1-200 gpurun_out/r2_bench_n1.json; tail -2 gpurun_out/r2_bench_n1.err
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err
grep -o '"value": [0-9.]*, "unit": "images/s", "n_gpus": 2, "steps": 20, "warmup": 5, "ms_per_step": [0-9.]*' gpurun_out/r2_bench_n2.json | head -1; tail -2 gpurun_out/r2_bench_n2.err
