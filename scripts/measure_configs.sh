#!/bin/bash
# Every BASELINE.json config + the reference's own runner.sh workloads through bench.py.
#   bash scripts/measure_configs.sh <arm: ours|reference> <N gpus> [steps] [warmup]     -> gpurun_out/configs_<arm>_n<N>.jsonl
# N = 1 runs in-process; N > 1 under torchrun (ours only; the reference is single-GPU by design and is launched once).
ARM=${1:-ours}; N=${2:-1}; STEPS=${3:-3}; WARM=${4:-3}
mkdir -p gpurun_out
OUT=gpurun_out/configs_${ARM}_n${N}.jsonl
: > $OUT
run() {   # name, bench args...
    name=$1; shift
    if [ "$N" -gt 1 ]; then
        LAUNCH="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300))"
    else
        LAUNCH="python"
    fi
    line=$(timeout 900 $LAUNCH bench.py --impl $ARM --gpus $N --steps $STEPS --warmup $WARM "$@" 2> gpurun_out/configs_${ARM}_n${N}_$name.err | tail -1)
    echo "{\"name\": \"$name\", \"result\": ${line:-null}}" >> $OUT
    python - "$name" "$line" <<'PY'
import json, sys
name, line = sys.argv[1], sys.argv[2]
try:
    d = json.loads(line)
    if "unavailable" in d:
        print(f"{name:28s} unavailable: {d['unavailable'][:90]}")
    else:
        e2e = d.get("e2e") or {}
        chk = d.get("agg_check") or {}
        print(f"{name:28s} {d['ms_per_step']:9.1f} ms/round  {d['value']:8.3f} rounds/s  e2e {e2e.get('value')}  agg_err {chk.get('agg_check_max_abs_err')} equal {chk.get('all_ranks_equal')}  fallbacks {d.get('library_fallbacks')}")
except Exception as ex:
    print(f"{name:28s} FAILED ({ex}): {line[:200]}")
PY
}
# ---- BASELINE.json configs (num_agents = 8 as named; on fewer GPUs the agents are time-multiplexed) -------------------------------
run fmnist_cnn_k8_rlr4      --model cnn_mnist --data fmnist  --train_size 60000 --agents 8 --theta 4 --num_corrupt 1 --poison_frac 0.5
run resnet18_k8_fedavg      --model resnet18  --data cifar10 --train_size 50000 --agents 8
run resnet18_k8_dba_rlr     --model resnet18  --data cifar10 --train_size 50000 --agents 8 --theta 4 --num_corrupt 2 --poison_frac 0.5
run vgg11_k8_comed          --model vgg11     --data cifar10 --train_size 50000 --agents 8 --aggr comed
# ---- the reference's own runner.sh workloads (src/runner.sh:12-38) ----------------------------------------------------------------
run readme_fmnist_10        --model cnn_mnist --data fmnist  --train_size 60000 --agents 10
run readme_fmnist_10_attack --model cnn_mnist --data fmnist  --train_size 60000 --agents 10 --num_corrupt 1 --poison_frac 0.5 --theta 4
run runner_cifar_40         --model cnn_cifar --data cifar10 --train_size 50000 --agents 40 --num_corrupt 4 --poison_frac 0.5 --theta 8
run runner_fedemnist_33     --model cnn_mnist --data fedemnist --train_size 676600 --agents 3383 --agent_frac 0.01 --num_corrupt 338 --poison_frac 0.5 --theta 8
cat $OUT | wc -l
