#!/bin/bash
# One gpurun --gpus 2 call of round 2 (kept for reproducibility): 2-rank checks (self-consistency + oracle) and all-reduce overlap A/B.
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tools/ddp_check.py > gpurun_out/ddp_check.log 2>&1; tail -2 gpurun_out/ddp_check.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 tests/ddp_oracle_check.py > gpurun_out/ddp_oracle_check.log 2>&1; tail -2 gpurun_out/ddp_oracle_check.log
python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-also > gpurun_out/b2_n1.json 2> gpurun_out/b2_n1.err; cut -c1-260 gpurun_out/b2_n1.json
python bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/b2_default.json 2> gpurun_out/b2_default.err; cut -c1-260 gpurun_out/b2_default.json; tail -2 gpurun_out/b2_default.err
python bench.py --gpus 2 --steps 20 --warmup 5 --comm-sms 0 > gpurun_out/b2_sms0.json 2> gpurun_out/b2_sms0.err; cut -c1-260 gpurun_out/b2_sms0.json
NCCL_MAX_CTAS=8 python bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/b2_cta8.json 2> gpurun_out/b2_cta8.err; cut -c1-260 gpurun_out/b2_cta8.json
NCCL_MAX_CTAS=16 python bench.py --gpus 2 --steps 20 --warmup 5 --comm-sms 16 > gpurun_out/b2_cta16.json 2> gpurun_out/b2_cta16.err; cut -c1-260 gpurun_out/b2_cta16.json
NCCL_MAX_CTAS=4 python bench.py --gpus 2 --steps 20 --warmup 5 --comm-sms 4 > gpurun_out/b2_cta4.json 2> gpurun_out/b2_cta4.err; cut -c1-260 gpurun_out/b2_cta4.json
