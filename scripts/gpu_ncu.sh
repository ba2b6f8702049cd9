#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"umma|fused_aggregate" -c 16 -f -o gpurun_out/prof_hot python scripts/profile_kernels.py all > gpurun_out/ncu_hot.log 2>&1
tail -3 gpurun_out/ncu_hot.log
ls -la gpurun_out/*.ncu-rep
timeout 600 python -m pytest tests/test_gpu_native.py -m gpu -q -x -k "stride2" 2>&1 | tail -5
