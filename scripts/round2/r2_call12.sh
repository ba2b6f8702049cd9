#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --durations=3 > gpurun_out/c12_pytest.txt 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/c12_pytest.txt | cut -c1-300
grep -E "FAILED|^E  " gpurun_out/c12_pytest.txt | head
timeout 900 python scripts/bench_aggregate.py 8,33,40,128,400 2>&1 | tee gpurun_out/c12_bench_aggregate.txt
for m in "cnn_mnist fmnist" "resnet18 cifar10"; do set -- $m
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/c12_launches_$1.csv python scripts/profile_step.py --trainer native --model $1 --data $2 --steps 3 > gpurun_out/c12_profile_$1.log 2>&1; tail -1 gpurun_out/c12_profile_$1.log
done
b() { name=$1; shift; env "$@" timeout 200 python bench.py --steps 3 --warmup 3 --no_e2e --model cnn_mnist --data fmnist --train_size 60000 --agents 10 > gpurun_out/c12_bench_$name.json 2> gpurun_out/c12_bench_$name.err; python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/c12_bench_{n}.json").read().strip().splitlines()[-1])
    print(f"bench fmnist-cnn-10 {n}: {d['ms_per_step']:.1f} ms/round fallbacks={d.get('library_fallbacks')}")
except Exception as e:
    print(f"bench {n}: FAILED {e}", open(f"gpurun_out/c12_bench_{n}.err").read()[-800:])
PY
}
b fused NONE=1
b unfused RLR_FUSE_DROPOUT=0
b fused2 NONE=1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"umma_|bn_|channel_reduce|fused_aggregate|sgd_step|gather_im2col" -c 60 -o gpurun_out/c12_ncu_hot python scripts/profile_kernels.py all > gpurun_out/c12_ncu_hot.log 2>&1; tail -2 gpurun_out/c12_ncu_hot.log
