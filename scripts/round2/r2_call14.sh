#!/bin/bash
mkdir -p gpurun_out
for fd in 1 0; do for nf in 1 2; do
  RLR_FUSE_DROPOUT=$fd timeout 120 python scripts/diag_handoff.py $nf 2>&1 | tail -2
done; done | tee gpurun_out/c14_diag_handoff.txt
for m in "cnn_mnist fmnist" "resnet18 cifar10"; do set -- $m
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/c14_launches_$1.csv python scripts/profile_step.py --trainer native --model $1 --data $2 --steps 3 > gpurun_out/c14_profile_$1.log 2>&1; tail -1 gpurun_out/c14_profile_$1.log
done
# ncu --set full of the hot kernels at HEAD; the report is exported to csv ON the box (the .ncu-rep itself is too large to bring back whole)
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"umma_|bn_|channel_reduce|fused_aggregate|sgd_step|gather_im2col" -c 60 -o /tmp/c14_ncu_hot python scripts/profile_kernels.py all > gpurun_out/c14_ncu_hot.log 2>&1; tail -2 gpurun_out/c14_ncu_hot.log
ncu -i /tmp/c14_ncu_hot.ncu-rep --page raw --csv > gpurun_out/c14_ncu_hot_raw.csv 2>/dev/null
ncu -i /tmp/c14_ncu_hot.ncu-rep --page details --csv > gpurun_out/c14_ncu_hot_details.csv 2>/dev/null
ncu -i /tmp/c14_ncu_hot.ncu-rep --page source --csv -k regex:umma_conv_gemm_kernel --launch-skip 0 --launch-count 1 > gpurun_out/c14_ncu_conv_source.csv 2>/dev/null
# a smaller report that fits: the generic conv forward + dgrad + wgrad at the layer-3 shape only
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"umma_" -c 6 -o gpurun_out/c14_ncu_conv python scripts/profile_kernels.py fwd > gpurun_out/c14_ncu_conv.log 2>&1
ls -la gpurun_out | head -20; du -sh gpurun_out
