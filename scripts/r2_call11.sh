#!/bin/bash
# all BASELINE.json configs + reference runner.sh workloads, both arms, 1 GPU
bash scripts/measure_configs.sh reference 1 2 1 2>&1 | tee gpurun_out/c11_configs_reference.txt
bash scripts/measure_configs.sh ours 1 3 3 2>&1 | tee gpurun_out/c11_configs_ours.txt
# agents in flight on the reference's README workload
for nf in 2 4; do
  timeout 300 python bench.py --model cnn_mnist --data fmnist --train_size 60000 --agents 10 --steps 3 --warmup 3 --no_e2e --agents_in_flight $nf 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('readme_fmnist_10 agents_in_flight=$nf: %.1f ms/round' % d['ms_per_step'])" | tee -a gpurun_out/c11_configs_ours.txt
done
