#!/bin/bash
# One gpurun call of round 2 (kept for reproducibility): GPU tests, bench A/B of the window operand staging, ncu launch list.
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --deselect tests/test_headline_gpu.py > gpurun_out/pytest_gpu.log 2>&1; tail -25 gpurun_out/pytest_gpu.log
python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-also > gpurun_out/bench_win.json 2> gpurun_out/bench_win.err; cut -c1-400 gpurun_out/bench_win.json; tail -3 gpurun_out/bench_win.err
B200_WINDOW=0 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-also > gpurun_out/bench_nowin.json 2> gpurun_out/bench_nowin.err; cut -c1-400 gpurun_out/bench_nowin.json
B200_WIN_SA=3 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-also > gpurun_out/bench_win_sa3.json 2> gpurun_out/bench_win_sa3.err; cut -c1-300 gpurun_out/bench_win_sa3.json
B200_WIN_TPB=2 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-also > gpurun_out/bench_win_tpb2.json 2> gpurun_out/bench_win_tpb2.err; cut -c1-300 gpurun_out/bench_win_tpb2.json
python bench.py --recipe inference --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_inf_win.json 2> gpurun_out/bench_inf_win.err; cut -c1-300 gpurun_out/bench_inf_win.json
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 1300 --csv --log-file gpurun_out/launches_win.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-graph --no-also > gpurun_out/ncu.log 2>&1
tail -2 gpurun_out/ncu.log | cut -c 1-300
