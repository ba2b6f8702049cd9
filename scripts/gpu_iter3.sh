#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_native.py -m gpu -q -x 2>&1 | tail -5 | cut -c1-300 | tee gpurun_out/iter_tests.txt
timeout 600 python bench.py --steps 3 --warmup 3 --trainer native --no_e2e > gpurun_out/bench_native.log 2>&1; tail -1 gpurun_out/bench_native.log | grep -o '"ms_per_step": [0-9.]*'
