#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_native.py -m gpu -q -x -k "stem_gemm or sm100_vs_fp32 or handoff or learns" > gpurun_out/c17_pytest_stem.txt 2>&1; echo "pytest stem rc=$?"; tail -3 gpurun_out/c17_pytest_stem.txt | cut -c1-300
grep -E "FAILED|^E  |Timeout" gpurun_out/c17_pytest_stem.txt | head
stem() { env "$@" python - <<'PY'
import os, torch
from rlr_b200 import ops
e = ops.ext()
for name, M, N, kv in (("resnet18 stem 3->64", 256 * 32 * 32, 64, 27), ("cnn_mnist conv1 1->32", 256 * 26 * 26, 32, 9), ("cnn_cifar conv1 3->64", 256 * 30 * 30, 64, 27)):
    A = torch.randn(M, 64, device="cuda").bfloat16(); W = torch.randn(N, kv, device="cuda").bfloat16(); out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    bias = torch.zeros(N, device="cuda")
    for _ in range(3): e.stem_gemm_bf16(A, W, out, bias, True, None, 0, 0, 0, None)
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g):
            for _ in range(20): e.stem_gemm_bf16(A, W, out, bias, True, None, 0, 0, 0, None)
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1000 / 20)
    print(f"RLR_STEM_PERSISTENT={os.environ.get('RLR_STEM_PERSISTENT', '1')} stem GEMM {name}: {best:.1f} us  ({(M * 64 + M * N) * 2 / best / 1e3:.0f} GB/s A read + out write)")
PY
}
stem RLR_STEM_PERSISTENT=0 2>&1 | tee gpurun_out/c17_stem_gemm.txt
stem RLR_STEM_PERSISTENT=1 2>&1 | tee -a gpurun_out/c17_stem_gemm.txt
b() {   # name, env, bench args...
    name=$1; envv=$2; shift 2
    env $envv timeout 300 python bench.py --steps 3 --warmup 3 --no_e2e "$@" > gpurun_out/c17_bench_$name.json 2> gpurun_out/c17_bench_$name.err
    python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/c17_bench_{n}.json").read().strip().splitlines()[-1])
    print(f"bench {n}: {d['ms_per_step']:.1f} ms/round fallbacks={d.get('library_fallbacks')}")
except Exception as e:
    print(f"bench {n}: FAILED {e}", open(f"gpurun_out/c17_bench_{n}.err").read()[-800:])
PY
}
F="--model cnn_mnist --data fmnist --train_size 60000 --agents 10"
C="--model cnn_cifar --data cifar10 --train_size 50000 --agents 40 --num_corrupt 4 --poison_frac 0.5 --theta 8"
b fmnist10_stem_persistent RLR_STEM_PERSISTENT=1 $F
b fmnist10_stem_old RLR_STEM_PERSISTENT=0 $F
b cifar40_stem_persistent RLR_STEM_PERSISTENT=1 $C
b cifar40_stem_old RLR_STEM_PERSISTENT=0 $C
b headline_stem_persistent RLR_STEM_PERSISTENT=1
b headline_stem_old RLR_STEM_PERSISTENT=0
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/c17_pytest_full.txt 2>&1; echo "pytest full rc=$?"; tail -2 gpurun_out/c17_pytest_full.txt | cut -c1-300
