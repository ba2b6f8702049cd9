#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4 | cut -c1-250 | tee gpurun_out/pytest_gpu_full.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1 | tee gpurun_out/smoke.txt
timeout 900 python bench.py > gpurun_out/bench_default.log 2>&1; tail -1 gpurun_out/bench_default.log | cut -c1-300
timeout 900 python bench.py --impl reference --gpus 8 --steps 1 --warmup 1 > gpurun_out/bench_ref_n8.log 2>&1; tail -1 gpurun_out/bench_ref_n8.log | cut -c1-300
