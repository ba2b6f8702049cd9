#!/bin/bash
# 2 GPUs: multi-GPU tests at HEAD (persistent stem GEMM behind the ready flags, rewritten hand-off test), headline + FMNIST-10 benches
mkdir -p gpurun_out
NCCL_DEBUG=WARN timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -q -s > gpurun_out/c18_pytest_multi.txt 2>&1; echo "multi tests rc=$?"; tail -6 gpurun_out/c18_pytest_multi.txt | cut -c1-300
grep -E "FAILED|^E  |Timeout|noise|fused" gpurun_out/c18_pytest_multi.txt | head -12 | cut -c1-250
run() { name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 "$@" > gpurun_out/c18_bench_$name.json 2> gpurun_out/c18_bench_$name.err
  tail -1 gpurun_out/c18_bench_$name.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench $name N=%d: %.1f ms/round, %.3f rounds/s, e2e %s, agg_check %s, handoff %s, in flight %s, phases %s' % (d['n_gpus'], d['ms_per_step'], d['value'], (d.get('e2e') or {}).get('value'), d.get('agg_check'), d['config'].get('fused_handoff'), d['config'].get('agents_in_flight_used'), d.get('phase_ms_per_round_rank0')))" || tail -5 gpurun_out/c18_bench_$name.err
}
run headline
run fmnist10 --model cnn_mnist --data fmnist --train_size 60000 --agents 10 --num_corrupt 1 --poison_frac 0.5 --theta 4
run cifar40 --model cnn_cifar --data cifar10 --train_size 50000 --agents 40 --num_corrupt 4 --poison_frac 0.5 --theta 8
