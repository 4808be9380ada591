#!/bin/bash
# One gpurun --gpus 2 call of round 2 (kept for reproducibility): gradient buckets on the wire in fp32 vs bf16.
mkdir -p gpurun_out
python bench.py --gpus 2 --steps 30 --warmup 5 > gpurun_out/b2_fp32.json 2> gpurun_out/b2_fp32.err; cut -c1-240 gpurun_out/b2_fp32.json
python bench.py --gpus 2 --steps 30 --warmup 5 --grad-wire bf16 > gpurun_out/b2_bf16.json 2> gpurun_out/b2_bf16.err; cut -c1-240 gpurun_out/b2_bf16.json; tail -2 gpurun_out/b2_bf16.err
B200_GRAD_WIRE=bf16 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 tests/ddp_oracle_check.py > gpurun_out/ddp_oracle_check_bf16.log 2>&1; tail -1 gpurun_out/ddp_oracle_check_bf16.log
python bench.py --gpus 2 --steps 30 --warmup 5 --grad-wire bf16 --bucket-mb 50 > gpurun_out/b2_bf16_b50.json 2> gpurun_out/b2_bf16_b50.err; cut -c1-240 gpurun_out/b2_bf16_b50.json
