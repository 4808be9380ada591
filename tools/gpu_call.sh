#!/bin/bash
# One gpurun call of round 2 (kept for reproducibility): GPU tests, per-layer timings, bench.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --deselect tests/test_headline_gpu.py > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
for st in 0 1; do
python tools/one_layer.py conv 64 128 64 128 10 $st
python tools/one_layer.py conv 64 64 128 256 10 $st
python tools/one_layer.py conv 64 32 256 512 10 $st
python tools/one_layer.py convt 64 64 256 64 10 $st
python tools/one_layer.py convt 64 32 512 128 10 $st
python tools/one_layer.py convt 64 16 1024 256 10 $st
done > gpurun_out/one_layer_cio.log 2>&1
cat gpurun_out/one_layer_cio.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-also > gpurun_out/bench_cio.json 2> gpurun_out/bench_cio.err; cut -c1-300 gpurun_out/bench_cio.json; tail -3 gpurun_out/bench_cio.err
B200_WINDOW=0 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-also > gpurun_out/bench_cio_nowin.json 2> gpurun_out/bench_cio_nowin.err; cut -c1-300 gpurun_out/bench_cio_nowin.json
python bench.py --recipe inference --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_inf_cio.json 2> gpurun_out/bench_inf_cio.err; cut -c1-300 gpurun_out/bench_inf_cio.json
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1300 --csv --log-file gpurun_out/launches_cio.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph --no-also > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log | cut -c 1-200
