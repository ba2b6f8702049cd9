#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/c20_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/c20_pytest.txt | cut -c1-300
grep -E "FAILED|^E  |Timeout" gpurun_out/c20_pytest.txt | head
python scripts/bench_small_gemm.py 2>&1 | tee gpurun_out/c20_small_gemm.txt
b() {   # name, env, bench args...
    name=$1; envv=$2; shift 2
    env $envv timeout 300 python bench.py --steps 3 --warmup 3 --no_e2e "$@" > gpurun_out/c20_bench_$name.json 2> gpurun_out/c20_bench_$name.err
    python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/c20_bench_{n}.json").read().strip().splitlines()[-1])
    print(f"bench {n}: {d['ms_per_step']:.1f} ms/round fallbacks={d.get('library_fallbacks')}")
except Exception as e:
    print(f"bench {n}: FAILED {e}", open(f"gpurun_out/c20_bench_{n}.err").read()[-800:])
PY
}
F="--model cnn_mnist --data fmnist --train_size 60000 --agents 10"
C="--model cnn_cifar --data cifar10 --train_size 50000 --agents 40 --num_corrupt 4 --poison_frac 0.5 --theta 8"
b fmnist10_relupool RLR_FUSE_RELU_POOL=1 $F
b fmnist10_relupool_off RLR_FUSE_RELU_POOL=0 $F
b cifar40_relupool RLR_FUSE_RELU_POOL=1 $C
b cifar40_relupool_off RLR_FUSE_RELU_POOL=0 $C
for m in "cnn_mnist fmnist" "cnn_cifar cifar10"; do set -- $m
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/c20_launches_$1.csv python scripts/profile_step.py --trainer native --model $1 --data $2 --steps 5 > gpurun_out/c20_profile_$1.log 2>&1; tail -1 gpurun_out/c20_profile_$1.log
done
