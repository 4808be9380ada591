#!/bin/bash
# One gpurun call of round 2 (kept for reproducibility): final evidence — GPU suite, smoke, bench lines (headline, reference arm, config 4), ncu launch list.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/r2_pytest_gpu.log 2>&1; tail -3 gpurun_out/r2_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; tail -1 gpurun_out/r2_smoke.log | cut -c1-300
python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; cut -c1-200 gpurun_out/r2_bench_n1.json; tail -2 gpurun_out/r2_bench_n1.err
python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err; cut -c1-200 gpurun_out/r2_bench_reference.json
python bench.py --recipe wgan_perceptual_style --steps 20 --warmup 5 > gpurun_out/r2_bench_config4_n1.json 2> gpurun_out/r2_bench_config4_n1.err; cut -c1-200 gpurun_out/r2_bench_config4_n1.json
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1300 --csv --log-file gpurun_out/r2_bench_launches_ncu.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph --no-also > gpurun_out/ncu.log 2>&1
tail -1 gpurun_out/ncu.log | cut -c1-120
