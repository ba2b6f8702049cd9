#!/bin/bash
# headline bench at N ranks (run with gpurun --gpus N)
set -x
N=${N:-2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps ${STEPS:-5} --warmup 3 > gpurun_out/bench_ours_n$N.log 2>&1; grep -v Warning gpurun_out/bench_ours_n$N.log | tail -2 | cut -c1-2800
