#!/bin/bash
# 2 GPUs at HEAD: multi-GPU tests, then the NVLink byte accounting
mkdir -p gpurun_out
NCCL_DEBUG=WARN timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -s > gpurun_out/c22_pytest_multi.txt 2>&1; echo "multi tests rc=$?"; tail -3 gpurun_out/c22_pytest_multi.txt | cut -c1-300
bash scripts/nvlink_count.sh 2>&1 | tail -8
