#!/bin/bash
# tcgen05 kernel bring-up: quick probes first (bounded), then the native-kernel test file, then the native bench
set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_native.py -m gpu -x -q -s -k "${K:-wgrad}" 2>&1 | tail -40 | tee gpurun_out/native_a.txt
timeout 600 python -m pytest tests/test_gpu_native.py -m gpu -q -s -k "native_net or native_trainer" 2>&1 | tail -40 | cut -c1-300 | tee gpurun_out/native_b.txt
timeout 600 python bench.py --steps 2 --warmup 3 --trainer native > gpurun_out/bench_native.log 2>&1; tail -5 gpurun_out/bench_native.log | cut -c1-3000
