"""Per-CTA timeline of the generic tcgen05 conv / GEMM kernel (globaltimer stamps written by the kernel when ops.ext().set_conv_trace
is armed): where does a CTA's life go -- set-up, first TMA round trip, main loop, epilogue -- and how are CTAs packed on the SMs.

    python scripts/trace_conv.py --layers l2,l3,l4 --dirs fwd,dgrad [--occ3 0]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from rlr_b200 import ops  # noqa: E402
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from bench_convs import LAYERS  # noqa: E402

DEV, BF = "cuda:0", torch.bfloat16
ap = argparse.ArgumentParser()
ap.add_argument("--layers", default="l2,l3,l4")
ap.add_argument("--dirs", default="fwd")
ap.add_argument("--occ3", type=int, default=1)
ap.add_argument("--split", type=int, default=0, help="1: two TMA producer threads per CTA")
ap.add_argument("--pair", type=int, default=0, help="0 single-CTA kernel | 1 CTA pairs | 2 pairs with the deep ring")
a = ap.parse_args()
e = ops.ext()
e.set_conv_occ3(a.occ3)
e.set_conv_2cta(a.pair)
e.set_conv_split_producer(bool(a.split))
for name in a.layers.split(","):
    B, H, Cin, Cout, k, s, p = LAYERS[name]
    Ho = (H + 2 * p - k) // s + 1
    x = torch.randn(B, H, H, Cin, device=DEV).to(BF)
    w = (torch.randn(Cout, k, k, Cin, device=DEV) * 0.05).to(BF)
    y = torch.empty(B, Ho, Ho, Cout, device=DEV, dtype=BF)
    dy = torch.randn(B, Ho, Ho, Cout, device=DEV).to(BF)
    dx = torch.empty_like(x)
    for d in a.dirs.split(","):
        fn = (lambda: ops.conv2d_fwd_sm100(x, w, None, y, s, p, False, None, tag=(name, "t"))) if d == "fwd" else \
             (lambda: ops.conv2d_dgrad_sm100(dy, w, dx, s, p, False))
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        buf = torch.zeros(8192 * 8, dtype=torch.int64, device=DEV)
        buf2 = torch.zeros(8192 * 8, dtype=torch.int64, device=DEV)
        e.set_conv_trace(buf)
        fn()
        e.set_conv_trace(buf2)        # a second, back-to-back launch: its first CTA start minus the first launch's last CTA end = launch gap
        fn()
        torch.cuda.synchronize()
        e.set_conv_trace(None)
        t2 = buf2.view(-1, 8).cpu()
        t2 = t2[t2[:, 0] > 0]
        t = buf.view(-1, 8).cpu()
        t = t[t[:, 0] > 0]
        n = t.shape[0]
        t0 = t[:, 0].min()
        rel = (t[:, :7] - t0).double() / 1e3          # us since the first CTA started
        span = float(rel[:, 6].max())
        q = lambda v: [round(float(x_), 2) for x_ in torch.quantile(v, torch.tensor([0.1, 0.5, 0.9], dtype=torch.float64))]
        gap = (float(t2[:, 0].min()) - float(t[:, 6].max())) / 1e3
        print(f"== {name} {d} pair={a.pair}: {n} CTAs, kernel span {span:.1f} us; gap to the next launch's first CTA {gap:.2f} us "
              f"(next launch: first CTA -> last CTA start {(float(t2[:, 0].max()) - float(t2[:, 0].min())) / 1e3:.2f} us)")
        print("   CTA start (10/50/90 %)        :", q(rel[:, 0]))
        print("   set-up  (entry -> barriers)    :", q(rel[:, 1] - rel[:, 0]))
        ld = t[:, 3] > 0                              # CTAs that issue MMAs (all of them, or the pair leaders)
        print("   setup done -> first data landed:", q((rel[:, 3] - rel[:, 1])[ld]))
        print("   main loop (first data -> issued):", q((rel[:, 4] - rel[:, 3])[ld]))
        print("   drain (issued -> acc complete) :", q((rel[:, 5] - rel[:, 4])[ld]))
        print("   epilogue: TMEM -> smem staging :", q(rel[:, 2] - rel[:, 5]))
        print("   epilogue: smem -> global stores:", q(rel[:, 6] - rel[:, 2]))
        print("   CTA life                       :", q(rel[:, 6] - rel[:, 0]))
        sm = t[:, 7]
        per_sm = torch.bincount(sm)
        print("   CTAs per SM: min/max", int(per_sm[per_sm > 0].min()), int(per_sm.max()), " SMs used", int((per_sm > 0).sum()))
        # one SM's timeline
        s0 = int(sm[0])
        mine = rel[sm == s0]
        order = torch.argsort(mine[:, 0])
        print(f"   SM {s0}: (start, first data, issued, acc done, end) per CTA")
        for i in order[:8]:
            r = mine[i]
            print("      ", [round(float(v), 1) for v in (r[0], r[3], r[4], r[5], r[6])])
        # note: in pair mode only the leader CTA of a pair has main-loop stamps (columns 3, 4); the peer's are 0
