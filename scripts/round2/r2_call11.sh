#!/bin/bash
# all BASELINE.json configs + reference runner.sh workloads, both arms, 1 GPU
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_native.py -m gpu -q -k "fused_dropout or learns or fp32_autograd" > gpurun_out/c11_pytest.txt 2>&1
rc=$?; echo "fused-dropout / net tests rc=$rc"; tail -3 gpurun_out/c11_pytest.txt | cut -c1-300; grep -E "^E  " gpurun_out/c11_pytest.txt | head -8
if [ $rc -ne 0 ]; then export RLR_FUSE_DROPOUT=0; echo "falling back to RLR_FUSE_DROPOUT=0 for the measurements"; fi
bash scripts/measure_configs.sh reference 1 2 1 2>&1 | tee gpurun_out/c11_configs_reference.txt
bash scripts/measure_configs.sh ours 1 3 3 2>&1 | tee gpurun_out/c11_configs_ours.txt
# agents in flight on the reference's README workload
for nf in 2 4; do
  timeout 300 python bench.py --model cnn_mnist --data fmnist --train_size 60000 --agents 10 --steps 3 --warmup 3 --no_e2e --agents_in_flight $nf 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('readme_fmnist_10 agents_in_flight=$nf: %.1f ms/round' % d['ms_per_step'])" | tee -a gpurun_out/c11_configs_ours.txt
done
timeout 600 python scripts/bench_aggregate.py 2>&1 | tee gpurun_out/c11_bench_aggregate.txt
