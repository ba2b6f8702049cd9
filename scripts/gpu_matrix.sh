#!/bin/bash
# CLI config matrix on one GPU: every aggregator / defence / attack / model family through federated.py, 2 rounds each.
mkdir -p gpurun_out
run() { echo "=== $*"; timeout 300 python federated.py --rounds 2 --local_ep 1 --bs 64 --synthetic 2048 --synthetic_val 256 --log_dir gpurun_out/logs --no_tensorboard "$@" 2>&1 | grep -E "Val_Loss|Poison Loss|Error|error|Traceback|finished" | tail -4; }
run --data=fmnist --num_agents=4 --num_corrupt=1 --poison_frac=0.5 --robustLR_threshold=3
run --data=fmnist --num_agents=4 --aggr=comed --pattern_type=copyright --num_corrupt=1 --poison_frac=0.5
run --data=fmnist --num_agents=6 --aggr=sign --server_lr=0.001 --agent_frac=0.5
run --data=fmnist --num_agents=3 --clip=0.5 --noise=0.001 --server_clip --diagnostics --num_corrupt=1 --poison_frac=0.5 --robustLR_threshold=2
run --data=cifar10 --num_agents=4 --num_corrupt=2 --poison_frac=0.5 --robustLR_threshold=3
run --data=cifar10 --model=resnet18 --num_agents=2 --checkpoint gpurun_out/ck.pt
run --data=cifar10 --model=resnet18 --num_agents=2 --resume gpurun_out/ck.pt --rounds 3
run --data=cifar10 --model=vgg11 --num_agents=2 --aggr=comed
run --data=fedemnist --num_agents=20 --agent_frac=0.25 --num_corrupt=2 --poison_frac=0.5 --pattern_type=square --robustLR_threshold=2
run --data=cifar10 --model=resnet18 --num_agents=2 --trainer=torch --dtype=fp32
run --data=cifar10 --model=cnn_cifar --num_agents=2 --no_graphs --profile_phases
