#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --durations=3 > gpurun_out/c16_pytest.txt 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/c16_pytest.txt | cut -c1-300
grep -E "FAILED|^E  |Timeout" gpurun_out/c16_pytest.txt | head
for cfg in "native 2" "torch 1" "torch 2"; do set -- $cfg
  timeout 150 python scripts/stress_inflight.py $1 $2 4 8 > gpurun_out/c16_stress_$1_$2.txt 2>&1; echo "stress $1 in-flight $2: rc=$? $(tail -1 gpurun_out/c16_stress_$1_$2.txt | cut -c1-200)"
done
b() {   # name, bench args...
    name=$1; shift
    timeout 300 python bench.py --steps 3 --warmup 3 --no_e2e "$@" > gpurun_out/c16_bench_$name.json 2> gpurun_out/c16_bench_$name.err
    python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/c16_bench_{n}.json").read().strip().splitlines()[-1])
    print(f"bench {n}: {d['ms_per_step']:.1f} ms/round fallbacks={d.get('library_fallbacks')}")
except Exception as e:
    print(f"bench {n}: FAILED {e}", open(f"gpurun_out/c16_bench_{n}.err").read()[-800:])
PY
}
b fmnist10 --model cnn_mnist --data fmnist --train_size 60000 --agents 10
b cifar40 --model cnn_cifar --data cifar10 --train_size 50000 --agents 40 --num_corrupt 4 --poison_frac 0.5 --theta 8
b headline
python - <<'PY' 2>&1 | tee gpurun_out/c16_stem_gemm.txt
import torch
from rlr_b200 import ops
e = ops.ext()
for name, M, N, kv in (("resnet18 stem 3->64", 256 * 32 * 32, 64, 27), ("cnn_mnist conv1 1->32", 256 * 26 * 26, 32, 9), ("cnn_cifar conv1 3->64", 256 * 30 * 30, 64, 27)):
    A = torch.randn(M, 64, device="cuda").bfloat16(); W = torch.randn(N, kv, device="cuda").bfloat16(); out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    bias = torch.zeros(N, device="cuda")
    for _ in range(3): e.stem_gemm_bf16(A, W, out, bias, True, None, 0, 0, 0, None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): e.stem_gemm_bf16(A, W, out, bias, True, None, 0, 0, 0, None)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1000 / 20
    print(f"stem GEMM {name}: {us:.1f} us  ({(M * 64 + M * N) * 2 / us / 1e3:.0f} GB/s of A read + out write)")
PY
du -sh gpurun_out
