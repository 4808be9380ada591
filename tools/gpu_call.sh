#!/bin/bash
# One gpurun call of round 2 (kept for reproducibility): hardware probe, GPU tests, parity table, bench lines, ncu launch list.
mkdir -p gpurun_out
./tools/umma_probe > gpurun_out/umma_probe.log 2>&1; tail -4 gpurun_out/umma_probe.log
python -m pytest tests -m gpu -x -q --deselect tests/test_headline_gpu.py > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
python -m tests.parity_report --out gpurun_out/parity_r2.json > gpurun_out/parity.log 2>&1; tail -6 gpurun_out/parity.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; tail -c 1500 gpurun_out/bench_a.json; tail -3 gpurun_out/bench_a.err
B200_KPB64=2 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-also > gpurun_out/bench_kpb2.json 2> gpurun_out/bench_kpb2.err
python bench.py --recipe inference --steps 5 --warmup 3 > gpurun_out/bench_inf.json 2> gpurun_out/bench_inf.err; tail -3 gpurun_out/bench_inf.err
python bench.py --recipe minimaxgan_rmse --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_mm.json 2> gpurun_out/bench_mm.err
python bench.py --recipe wgan_d_only --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_donly.json 2> gpurun_out/bench_donly.err
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph --no-also > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log | cut -c 1-300
