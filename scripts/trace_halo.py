"""Per-tile clock64 timeline of CTA 0 of the halo conv kernel (tuning aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rlr_b200 import ops
DEV, BF = "cuda:0", torch.bfloat16
x = torch.randn(256, 32, 32, 64, device=DEV).to(BF)
w = (torch.randn(64, 576, device=DEV) * 0.05).to(BF)
y = torch.empty(256, 32, 32, 64, device=DEV, dtype=BF)
stats = torch.zeros(ops.STAT_SLOTS, 2, 64, device=DEV)
for mode in ("plain", "stats", "acc"):
    dbg = torch.zeros(16 * 8, dtype=torch.int64, device=DEV)
    for _ in range(2):
        ops.ext().conv3x3_halo_bf16(x, w, y, None, False, mode == "acc", stats if mode == "stats" else None, 0, dbg)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.ext().conv3x3_halo_bf16(x, w, y, None, False, mode == "acc", stats if mode == "stats" else None, 0, None)
    e1.record(); torch.cuda.synchronize()
    d = dbg.view(16, 8).cpu()
    t0 = int(d[0, 0])
    print(mode, "avg us/launch (back-to-back, warm L2):", e0.elapsed_time(e1) * 100)
    print(" tile: mma[start waitAcc waitHalo issued] epi[start gotAcc ldDone stored]  (cycles rel. to first)")
    for i in range(14):
        print("  %2d" % i, [int(v) - t0 for v in d[i]])
