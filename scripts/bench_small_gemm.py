"""Kernel time of the small dense layers of the reference CNNs on the tcgen05 GEMM (CUDA graph of 20 launches, best of 3):
which epilogue feature costs what.  Usage: python scripts/bench_small_gemm.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from rlr_b200 import ops  # noqa: E402

e = ops.ext()
DEV = "cuda:0"
step = torch.zeros(1, dtype=torch.int64, device=DEV)


def timed(fn):
    for _ in range(3):
        fn()
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        with torch.cuda.graph(g):
            for _ in range(20):
                fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1000 / 20)
    return best


for M, N, K in ((256, 256, 128), (256, 128, 256), (256, 1024, 128), (256, 128, 1024)):
    A = torch.randn(M, K, device=DEV).bfloat16(); W = torch.randn(N, K, device=DEV).bfloat16()
    out = torch.empty(M, N, device=DEV, dtype=torch.bfloat16); bias = torch.randn(N, device=DEV)
    r = {}
    r["plain"] = timed(lambda: e.gemm_bf16(A, W, out, None, False, False, None))
    r["bias"] = timed(lambda: e.gemm_bf16(A, W, out, bias, False, False, None))
    r["bias+relu"] = timed(lambda: e.gemm_bf16(A, W, out, bias, True, False, None))
    r["bias+relu+drop"] = timed(lambda: e.gemm_bf16(A, W, out, bias, True, False, None, drop_p=0.5, drop_seed=3, drop_step=step, drop_stream=7))
    r["accumulate"] = timed(lambda: e.gemm_bf16(A, W, out, None, False, True, None))
    print(f"gemm {M}x{N}x{K}: " + "  ".join(f"{k} {v:.1f} us" for k, v in r.items()), flush=True)
