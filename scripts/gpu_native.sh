#!/bin/bash
# tcgen05 kernel bring-up: quick GEMM probe first (bounded), then the native-kernel test file
set -x
mkdir -p gpurun_out
timeout 180 python -m pytest tests/test_gpu_native.py -m gpu -x -q -k "gemm" 2>&1 | tail -25 | tee gpurun_out/native_gemm.txt
timeout 600 python -m pytest tests/test_gpu_native.py -m gpu -q -k "not gemm" 2>&1 | tail -60 | tee gpurun_out/native_rest.txt
