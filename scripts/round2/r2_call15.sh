#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --durations=3 > gpurun_out/c15_pytest.txt 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/c15_pytest.txt | cut -c1-300
grep -E "FAILED|^E  " gpurun_out/c15_pytest.txt | head
for i in 1 2 3; do timeout 200 python -m pytest tests/test_gpu_native.py -m gpu -q -s -k "fused_handoff_equals" 2>&1 | grep -E "round 1|passed|failed" | cut -c1-200; done | tee gpurun_out/c15_handoff_repeat.txt
b() {   # name, env assignment, bench args...
    name=$1; envv=$2; shift 2
    env $envv timeout 300 python bench.py --steps 3 --warmup 3 --no_e2e "$@" > gpurun_out/c15_bench_$name.json 2> gpurun_out/c15_bench_$name.err
    python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/c15_bench_{n}.json").read().strip().splitlines()[-1])
    print(f"bench {n}: {d['ms_per_step']:.1f} ms/round fallbacks={d.get('library_fallbacks')}")
except Exception as e:
    print(f"bench {n}: FAILED {e}", open(f"gpurun_out/c15_bench_{n}.err").read()[-800:])
PY
}
F="--model cnn_mnist --data fmnist --train_size 60000 --agents 10"
C="--model cnn_cifar --data cifar10 --train_size 50000 --agents 40 --num_corrupt 4 --poison_frac 0.5 --theta 8"
b fmnist10_halo_any RLR_HALO_ANY=1 $F
b fmnist10_halo_old RLR_HALO_ANY=0 $F
b fmnist10_halo_any2 RLR_HALO_ANY=1 $F
b cifar40_halo_any RLR_HALO_ANY=1 $C
b cifar40_halo_old RLR_HALO_ANY=0 $C
for m in "cnn_mnist fmnist" "cnn_cifar cifar10" "resnet18 cifar10"; do set -- $m
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none --csv --log-file gpurun_out/c15_launches_$1.csv python scripts/profile_step.py --trainer native --model $1 --data $2 --steps 5 > gpurun_out/c15_profile_$1.log 2>&1; tail -1 gpurun_out/c15_profile_$1.log
done
# ncu --set full of the hot kernels at HEAD, caches left warm (as inside the training graph); exported to csv ON the box
timeout 600 ncu --set full --clock-control none --cache-control none --import-source on -k regex:"umma_|bn_|channel_reduce|fused_aggregate|sgd_step|gather_im2col" -c 60 -o /tmp/c15_ncu_hot python scripts/profile_kernels.py all > gpurun_out/c15_ncu_hot.log 2>&1; tail -2 gpurun_out/c15_ncu_hot.log
ncu -i /tmp/c15_ncu_hot.ncu-rep --page raw --csv > gpurun_out/c15_ncu_hot_raw.csv 2>/dev/null
ncu -i /tmp/c15_ncu_hot.ncu-rep --page details --csv > gpurun_out/c15_ncu_hot_details.csv 2>/dev/null
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/c15_bench_headline.json 2> gpurun_out/c15_bench_headline.err; tail -1 gpurun_out/c15_bench_headline.json | cut -c1-400
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
du -sh gpurun_out
