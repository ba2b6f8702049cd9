#!/bin/bash
# full GPU validation: all gpu-marked tests, smoke, headline bench (auto trainer)
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q --durations=8 2>&1 | tail -40 | cut -c1-250 | tee gpurun_out/pytest_gpu_full.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/smoke.txt
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_ours_n1.log 2>&1; tail -1 gpurun_out/bench_ours_n1.log | cut -c1-2600
