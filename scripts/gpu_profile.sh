#!/bin/bash
# launch lists (device time per kernel) for one eager local step of each trainer + native net tests with full output
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_native.py -m gpu -q -s -k "native_net or cnn_cifar" > gpurun_out/native_b_full.txt 2>&1; tail -5 gpurun_out/native_b_full.txt
for T in native torch; do
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_$T.csv python scripts/profile_step.py --trainer $T --steps 3 > gpurun_out/profile_$T.log 2>&1
  tail -2 gpurun_out/profile_$T.log
done
