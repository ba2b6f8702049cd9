#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/c23_pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/c23_pytest.txt | cut -c1-300
grep -E "FAILED|^E  |Timeout" gpurun_out/c23_pytest.txt | head
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
    print(f"RLR_STEM_PERSISTENT={os.environ.get('RLR_STEM_PERSISTENT', '2')} stem GEMM {name}: {best:.1f} us  ({(M * 64 + M * N) * 2 / best / 1e3:.0f} GB/s A read + out write)")
PY
}
for v in 0 1 2; do stem RLR_STEM_PERSISTENT=$v 2>&1 | tee -a gpurun_out/c23_stem_gemm.txt; done
b() {   # name, env, bench args...
    name=$1; envv=$2; shift 2
    env $envv timeout 300 python bench.py --steps 3 --warmup 3 --no_e2e "$@" > gpurun_out/c23_bench_$name.json 2> gpurun_out/c23_bench_$name.err
    python - "$name" <<'PY'
import json, sys
n = sys.argv[1]
try:
    d = json.loads(open(f"gpurun_out/c23_bench_{n}.json").read().strip().splitlines()[-1])
    print(f"bench {n}: {d['ms_per_step']:.1f} ms/round fallbacks={d.get('library_fallbacks')}")
except Exception as e:
    print(f"bench {n}: FAILED {e}", open(f"gpurun_out/c23_bench_{n}.err").read()[-800:])
PY
}
F="--model cnn_mnist --data fmnist --train_size 60000 --agents 10"
C="--model cnn_cifar --data cifar10 --train_size 50000 --agents 40 --num_corrupt 4 --poison_frac 0.5 --theta 8"
b fmnist10_stem2 RLR_STEM_PERSISTENT=2 $F
b fmnist10_stem1 RLR_STEM_PERSISTENT=1 $F
b cifar40_stem2 RLR_STEM_PERSISTENT=2 $C
b cifar40_stem1 RLR_STEM_PERSISTENT=1 $C
b headline_stats1 RLR_EPILOGUE_BN_STATS=1
b headline_stats0 RLR_EPILOGUE_BN_STATS=0
b headline_stats1b RLR_EPILOGUE_BN_STATS=1
b vgg11_stats1 RLR_EPILOGUE_BN_STATS=1 --model vgg11 --agents 8 --aggr comed
b vgg11_stats0 RLR_EPILOGUE_BN_STATS=0 --model vgg11 --agents 8 --aggr comed
