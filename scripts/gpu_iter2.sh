#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_native.py -m gpu -q -x 2>&1 | tail -15 | cut -c1-300 | tee gpurun_out/iter_tests.txt
timeout 600 python bench.py --steps 2 --warmup 3 --trainer native --no_e2e > gpurun_out/bench_native.log 2>&1; tail -1 gpurun_out/bench_native.log | grep -o '"ms_per_step": [0-9.]*'
RLR_WG_WAVES=1 timeout 600 python bench.py --steps 2 --warmup 3 --trainer native --no_e2e > gpurun_out/bench_native_w1.log 2>&1; tail -1 gpurun_out/bench_native_w1.log | grep -o '"ms_per_step": [0-9.]*'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_native.csv python scripts/profile_step.py --trainer native --steps 3 > gpurun_out/profile_native.log 2>&1
tail -1 gpurun_out/profile_native.log
