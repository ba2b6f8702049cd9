#!/bin/bash
# 2 GPUs, final: multi-GPU tests + headline and FMNIST-10 benches at HEAD (branch streams, two agents in flight, hand-off)
mkdir -p gpurun_out
NCCL_DEBUG=WARN timeout 500 python -m pytest tests/test_gpu_multi.py -m gpu -q -s > gpurun_out/c29_pytest_multi.txt 2>&1; echo "multi tests rc=$?"; tail -2 gpurun_out/c29_pytest_multi.txt | cut -c1-300
run() { name=$1; shift
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 3 --warmup 3 "$@" > gpurun_out/c29_bench_$name.json 2> gpurun_out/c29_bench_$name.err
  tail -1 gpurun_out/c29_bench_$name.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('bench $name N=%d: %.1f ms/round, %.3f rounds/s, e2e %s, agg_check %s, handoff %s, in flight %s' % (d['n_gpus'], d['ms_per_step'], d['value'], (d.get('e2e') or {}).get('value'), {k: (d.get('agg_check') or {}).get(k) for k in ('agg_check_max_abs_err','all_ranks_equal','flipped_kernel','flipped_oracle')}, d['config'].get('fused_handoff'), d['config'].get('agents_in_flight_used')))" || tail -5 gpurun_out/c29_bench_$name.err
}
run headline
run resnet18_k8 --agents 8 --theta 4 --num_corrupt 2 --poison_frac 0.5
run fmnist10 --model cnn_mnist --data fmnist --train_size 60000 --agents 10 --num_corrupt 1 --poison_frac 0.5 --theta 4
