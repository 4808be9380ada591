#!/bin/bash
# One gpurun call of round 2 (kept for reproducibility): 4-phase window mode (WIN 4) validation and A/B.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --deselect tests/test_headline_gpu.py > gpurun_out/pytest_gpu.log 2>&1; tail -6 gpurun_out/pytest_gpu.log
for w4 in 1 0; do
B200_WIN4=$w4 python tools/one_layer.py convt 64 64 256 64 10 0
B200_WIN4=$w4 python tools/one_layer.py convt 64 64 256 64 10 1
B200_WIN4=$w4 python tools/one_layer.py convt 64 64 128 64 10 0
B200_WIN4=$w4 python tools/one_layer.py convt 128 128 256 64 5 0
done > gpurun_out/one_layer_w4.log 2>&1; cat gpurun_out/one_layer_w4.log
python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-also > gpurun_out/bench_w4.json 2> gpurun_out/bench_w4.err; cut -c1-200 gpurun_out/bench_w4.json; tail -2 gpurun_out/bench_w4.err
B200_WIN4=0 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-also > gpurun_out/bench_w4off.json 2> gpurun_out/bench_w4off.err; cut -c1-200 gpurun_out/bench_w4off.json
python bench.py --recipe inference --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_inf_w4.json 2> gpurun_out/bench_inf_w4.err; cut -c1-200 gpurun_out/bench_inf_w4.json
