"""Single-GPU timing of the fused server-step kernel (ops/csrc/aggregate.cu) for real participant counts:
K in {8, 33, 40, 128, 400} x {avg, comed, sign}, n = ResNet-18 sized flat vector.  CUDA events around REPS back-to-back launches.
Algorithmic bytes = (K + 1) * n * 4 read + n * 6 written (fp32 + bf16 shadow); fraction of the MEASURED copy bandwidth (MEASURED_PEAKS.json)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from rlr_b200 import ops  # noqa: E402

DEV = "cuda:0"
n = 11190272
try:
    hbm = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:  # noqa: BLE001
    hbm = 6650.0
torch.manual_seed(0)
g = torch.randn(n, device=DEV)
out = torch.empty_like(g)
sh = torch.empty(n, device=DEV, dtype=torch.bfloat16)
flipped = torch.zeros(1, dtype=torch.int64, device=DEV)
rows = []
Ks = [int(k) for k in (sys.argv[1].split(",") if len(sys.argv) > 1 else "8,33,40,128,400".split(","))]
for K in Ks:
    ws = [g + 0.01 * torch.randn(n, device=DEV) for _ in range(K)]
    wt = [1.0 + (k % 3) for k in range(K)]
    for mode in ("avg", "comed", "sign"):
        reps = 5 if K <= 40 else 2
        fn = lambda: ops.fused_aggregate(g, ws, wt, mode, max(1, K // 4), 1.0, n_vote=n - 12288, out=out, out_bf16=sh, flipped=flipped)
        fn(); fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(3):           # plain launches (the wrapper builds a small pointer table per call; 30 us of host work vs >= 100 us kernels)
            e0.record()
            for _ in range(reps):
                fn()
            e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e3 / reps)
        gb = ((K + 1) * n * 4 + n * 6) / 1e9
        bw = gb / (best * 1e-6)
        rows.append(dict(K=K, mode=mode, us=best, gb=gb, gbs=bw, frac=bw / hbm))
        print(f"K={K:4d} {mode:5s} {best:9.1f} us  {gb:6.2f} GB  {bw:7.0f} GB/s  {bw / hbm:5.2f} of measured copy bandwidth ({hbm:.0f} GB/s)", flush=True)
    del ws
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rows, open("gpurun_out/bench_aggregate.json", "w"))
