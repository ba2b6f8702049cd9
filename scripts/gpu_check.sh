#!/bin/bash
# One gpurun call: GPU tests, smoke, headline bench (ours + reference).  Outputs land in gpurun_out/.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/pytest_gpu.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.txt
timeout 900 python bench.py --steps ${STEPS:-2} --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench_ours.txt
timeout 900 python bench.py --impl reference --steps 1 --warmup 1 2>&1 | tail -3 | tee gpurun_out/bench_ref.txt
