#!/bin/bash
# One gpurun call of round 2 (kept for reproducibility): conv_c1_fwd instruction diet — tests, bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_kernels_gpu.py tests/test_steps_gpu.py tests/test_networks_gpu.py -q -x > gpurun_out/r2_tail_tests.log 2>&1; echo rc=$?; tail -3 gpurun_out/r2_tail_tests.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-also > gpurun_out/r2_bench_tail.json 2> gpurun_out/r2_bench_tail.err; cut -c1-260 gpurun_out/r2_bench_tail.json; tail -2 gpurun_out/r2_bench_tail.err
python - <<'P'
import json
d = json.loads([l for l in open("gpurun_out/r2_bench_tail.json").read().splitlines() if l.startswith("{")][-1])
for k, v in sorted(d["tail"]["kernels"].items(), key=lambda kv: -kv[1]["ms"])[:8]:
    print(k, v)
print(d["clocks"], d["tail"]["ms_per_step"])
P
