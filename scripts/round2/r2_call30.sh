#!/bin/bash
mkdir -p gpurun_out
for m in "resnet18 cifar10" "cnn_mnist fmnist" "cnn_cifar cifar10"; do set -- $m
  timeout 70 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/c30_launches_$1.csv python scripts/profile_step.py --trainer native --model $1 --data $2 --steps 5 > gpurun_out/c30_profile_$1.log 2>&1; tail -1 gpurun_out/c30_profile_$1.log
done
