#!/bin/bash
# One gpurun call of round 2 (kept for reproducibility): smoke() and the default bench line at HEAD.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -1 gpurun_out/r2_smoke.log | cut -c1-200
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; cut -c1-200 gpurun_out/r2_bench_n1.json; tail -2 gpurun_out/r2_bench_n1.err
