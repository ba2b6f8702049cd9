#!/bin/bash
# Canned experiments, mirroring the reference launcher (src/runner.sh:12-38) on the B200 engine.
# Usage: bash scripts/runner.sh [NGPUS]   (NGPUS > 1 launches one rank per GPU with torchrun)
set -e
NG=${1:-1}
cd "$(dirname "$0")/.."
if [ "$NG" -gt 1 ]; then
  RUN="python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29600 federated.py"
else
  RUN="python federated.py"
fi
echo 'Calling scripts!'
rm -rf logs   # like the reference (src/runner.sh:6); comment out to keep old logs

# FMNIST: FedAvg / backdoor attack / attack + Robust LR (reference src/runner.sh:12-18)
$RUN --data=fmnist --local_ep=2 --bs=256 --num_agents=10 --rounds=200
$RUN --data=fmnist --local_ep=2 --bs=256 --num_agents=10 --rounds=200 --num_corrupt=1 --poison_frac=0.5
$RUN --data=fmnist --local_ep=2 --bs=256 --num_agents=10 --rounds=200 --num_corrupt=1 --poison_frac=0.5 --robustLR_threshold=4

# CIFAR-10: 40 agents, distributed backdoor attack (src/runner.sh:23-28)
$RUN --data=cifar10 --local_ep=2 --bs=256 --num_agents=40 --rounds=200
$RUN --data=cifar10 --local_ep=2 --bs=256 --num_agents=40 --rounds=200 --num_corrupt=4 --poison_frac=0.5
$RUN --data=cifar10 --local_ep=2 --bs=256 --num_agents=40 --rounds=200 --num_corrupt=4 --poison_frac=0.5 --robustLR_threshold=8

# Fed-EMNIST (non-IID, 3383 clients, 1% participation; needs ../data/Fed_EMNIST) (src/runner.sh:34-38)
$RUN --data=fedemnist --num_agents=3383 --agent_frac=0.01 --local_ep=10 --bs=64 --rounds=500 --snap=5
$RUN --data=fedemnist --num_agents=3383 --agent_frac=0.01 --num_corrupt=338 --poison_frac=0.5 --local_ep=10 --bs=64 --rounds=500 --snap=5
$RUN --data=fedemnist --num_agents=3383 --agent_frac=0.01 --num_corrupt=338 --poison_frac=0.5 --robustLR_threshold=8 --local_ep=10 --bs=64 --rounds=500 --snap=5

# B200-only extras: ResNet-18 / VGG-11 (BASELINE.json configs)
$RUN --data=cifar10 --model=resnet18 --local_ep=2 --bs=256 --num_agents=8 --rounds=100 --num_corrupt=2 --poison_frac=0.5 --robustLR_threshold=4
$RUN --data=cifar10 --model=vgg11 --aggr=comed --local_ep=2 --bs=256 --num_agents=8 --rounds=100
echo 'All experiments are finished!'
