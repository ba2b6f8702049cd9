#!/bin/bash
# compute-sanitizer pass over the kernel tests.  memcheck on the tensor-core / TMA kernels and the fused
# aggregation kernel; racecheck on the elementwise / reduction kernels.  Slow (10-50x): run on demand via gpurun.
set -x
mkdir -p gpurun_out
K=${K:-"gemm or halo or conv_fwd or wgrad"}
timeout ${T:-600} compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_native.py -m gpu -q -x -k "$K" 2>&1 | tail -25 | cut -c1-300 | tee gpurun_out/sanitize_memcheck.txt
timeout ${T:-600} compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "fused_aggregate or flat_sgd or softmax or round_init" 2>&1 | tail -15 | cut -c1-300 | tee gpurun_out/sanitize_racecheck.txt
