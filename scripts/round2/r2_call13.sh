#!/bin/bash
mkdir -p gpurun_out
for fd in 1 0; do for nf in 1 2; do
  RLR_FUSE_DROPOUT=$fd timeout 120 python scripts/diag_handoff.py $nf 2>&1 | tail -2
done; done | tee gpurun_out/c13_diag_handoff.txt
for m in "cnn_mnist fmnist" "resnet18 cifar10"; do set -- $m
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/c13_launches_$1.csv python scripts/profile_step.py --trainer native --model $1 --data $2 --steps 3 > gpurun_out/c13_profile_$1.log 2>&1; tail -1 gpurun_out/c13_profile_$1.log
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"umma_|bn_|channel_reduce|fused_aggregate|sgd_step|gather_im2col" -c 60 -o gpurun_out/c13_ncu_hot python scripts/profile_kernels.py all > gpurun_out/c13_ncu_hot.log 2>&1; tail -2 gpurun_out/c13_ncu_hot.log
ls -la gpurun_out | head -20; du -sh gpurun_out
