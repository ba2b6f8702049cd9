#!/bin/bash
# end-of-round validation on 1 GPU: all GPU tests, smoke, both bench arms, secondary configs, launch list, ncu of the conv kernels
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -6 | cut -c1-250 | tee gpurun_out/pytest_gpu_full.txt
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/smoke.txt
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_ours_n1.log 2>&1; tail -1 gpurun_out/bench_ours_n1.log | cut -c1-400
timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_ref_n1.log 2>&1; tail -1 gpurun_out/bench_ref_n1.log | cut -c1-300
timeout 600 python bench.py --steps 3 --warmup 3 --model cnn_mnist --data fmnist --train_size 60000 --theta 4 --no_e2e > gpurun_out/bench_fmnist_cnn.log 2>&1; tail -1 gpurun_out/bench_fmnist_cnn.log | cut -c1-300
timeout 600 python bench.py --steps 3 --warmup 3 --model vgg11 --aggr comed --no_e2e > gpurun_out/bench_vgg11_comed.log 2>&1; tail -1 gpurun_out/bench_vgg11_comed.log | cut -c1-300
timeout 600 python bench.py --steps 3 --warmup 3 --trainer torch --no_e2e > gpurun_out/bench_torchtrainer.log 2>&1; tail -1 gpurun_out/bench_torchtrainer.log | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_native.csv python scripts/profile_step.py --trainer native --steps 3 > gpurun_out/profile_native.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"umma" -c 12 -f -o gpurun_out/prof_conv python scripts/profile_kernels.py all > gpurun_out/ncu_conv.log 2>&1; tail -1 gpurun_out/ncu_conv.log
