#!/bin/bash
set -x
mkdir -p gpurun_out
timeout ${T:-300} python -m pytest tests/test_gpu_native.py -m gpu -q -s -x -k "$K" 2>&1 | tail -30 | cut -c1-400 | tee gpurun_out/one.txt
