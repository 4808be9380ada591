#!/bin/bash
# One gpurun call of round 2 (kept for reproducibility): GPU tests (incl. the input pipeline), PDL A/B.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --deselect tests/test_headline_gpu.py > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-also > gpurun_out/bench_pdl0.json 2> gpurun_out/bench_pdl0.err; cut -c1-260 gpurun_out/bench_pdl0.json
B200_PDL=2 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-also > gpurun_out/bench_pdl2.json 2> gpurun_out/bench_pdl2.err; cut -c1-260 gpurun_out/bench_pdl2.json
B200_PDL=1 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-also > gpurun_out/bench_pdl1.json 2> gpurun_out/bench_pdl1.err; cut -c1-260 gpurun_out/bench_pdl1.json
