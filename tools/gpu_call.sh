#!/bin/bash
# One gpurun call of round 2 (kept for reproducibility): full GPU suite incl. the headline-shape parity tests, parity table, bench, launch list.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu_full.log 2>&1; tail -5 gpurun_out/pytest_gpu_full.log
python -m tests.parity_report --out gpurun_out/parity_r2.json > gpurun_out/parity.log 2>&1; tail -2 gpurun_out/parity.log | cut -c1-300
python bench.py --steps 30 --warmup 5 > gpurun_out/bench_pf.json 2> gpurun_out/bench_pf.err; cut -c1-260 gpurun_out/bench_pf.json; tail -2 gpurun_out/bench_pf.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1300 --csv --log-file gpurun_out/launches_pf.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph --no-also > gpurun_out/ncu.log 2>&1
tail -1 gpurun_out/ncu.log | cut -c 1-200
