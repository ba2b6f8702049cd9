#!/bin/bash
set -x
N=${N:-8}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
NCCL_DEBUG=WARN timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -x -q -s 2>&1 | tail -15 | cut -c1-300 | tee gpurun_out/pytest_multi_n$N.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_ours_n$N.log 2>&1; grep -v Warning gpurun_out/bench_ours_n$N.log | tail -2 | cut -c1-2600
