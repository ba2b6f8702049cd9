#!/bin/bash
# 8-GPU run: headline bench (with e2e + agg_check) and the BASELINE.json / runner.sh configs on the fused P2P / NVLS path
N=${N:-8}
mkdir -p gpurun_out
OUT=gpurun_out/configs_ours_n${N}.jsonl
: > $OUT
run() {   # name, bench args...
    name=$1; shift
    line=$(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) \
           bench.py --gpus $N --steps 3 --warmup 3 "$@" 2> gpurun_out/n${N}_$name.err | tail -1)
    echo "{\"name\": \"$name\", \"result\": ${line:-null}}" >> $OUT
    python - "$name" "$line" <<'PY'
import json, sys
name, line = sys.argv[1], sys.argv[2]
try:
    d = json.loads(line)
    e2e = d.get("e2e") or {}
    chk = d.get("agg_check") or {}
    print(f"{name:26s} {d['ms_per_step']:8.1f} ms/round {d['value']:8.3f} rounds/s e2e {e2e.get('value')} agg_err {chk.get('agg_check_max_abs_err')} equal {chk.get('all_ranks_equal')} "
          f"flip {chk.get('flipped_kernel')}/{chk.get('flipped_oracle')} mc {d['config'].get('multicast')} handoff {d['config'].get('fused_handoff')} phases {d.get('phase_ms_per_round_rank0')} clocks {d.get('clocks', {}).get('sm_mhz')}")
except Exception as ex:
    print(f"{name:26s} FAILED ({ex}): {line[:300]}")
PY
}
run resnet18_k8_fedavg
run resnet18_k8_nohandoff  --no_e2e --no_fused_handoff
run resnet18_k8_dba_rlr    --no_e2e --theta 4 --num_corrupt 2 --poison_frac 0.5
run vgg11_k8_comed         --no_e2e --model vgg11 --aggr comed
run fmnist_cnn_k8_rlr4     --no_e2e --model cnn_mnist --data fmnist --train_size 60000 --theta 4 --num_corrupt 1 --poison_frac 0.5
run runner_cifar_40        --no_e2e --model cnn_cifar --agents 40 --num_corrupt 4 --poison_frac 0.5 --theta 8
run readme_fmnist_10       --no_e2e --model cnn_mnist --data fmnist --train_size 60000 --agents 10
