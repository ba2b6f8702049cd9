"""Per-layer timing of the tcgen05 convolution kernels on the ResNet-18 / VGG shapes (CUDA events around a captured graph of
REPS back-to-back launches; warm L2 like inside the training step).  One line per (layer, direction, variant):

    python scripts/bench_convs.py [--variants default,pair,...] [--layers l1,l2,l3,l4] [--reps 20]

Variants toggle the launcher knobs at run time (ops.ext().set_conv_2cta / set_conv_occ3 / ...) or module flags of ops/nn.py, so
one process measures them on the same box.  FLOPs = 2*B*H*W*Cin*Cout*k*k; fractions are of MEASURED_PEAKS.json (cuBLAS bf16)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from rlr_b200 import ops  # noqa: E402
from rlr_b200.ops import nn  # noqa: E402

DEV, BF = "cuda:0", torch.bfloat16
LAYERS = {  # name: (B, H, Cin, Cout, k, stride, pad)
    "l1": (256, 32, 64, 64, 3, 1, 1), "l2": (256, 16, 128, 128, 3, 1, 1), "l3": (256, 8, 256, 256, 3, 1, 1),
    "l4": (256, 4, 512, 512, 3, 1, 1),
    "l2s": (256, 32, 64, 128, 3, 2, 1), "l3s": (256, 16, 128, 256, 3, 2, 1), "l4s": (256, 8, 256, 512, 3, 2, 1),
    "l2d": (256, 32, 64, 128, 1, 2, 0), "l3d": (256, 16, 128, 256, 1, 2, 0), "l4d": (256, 8, 256, 512, 1, 2, 0),
    "v1": (256, 16, 64, 128, 3, 1, 1), "v2": (256, 8, 128, 256, 3, 1, 1), "v3": (256, 4, 256, 512, 3, 1, 1), "v4": (256, 2, 512, 512, 3, 1, 1),
}


def peak():
    try:
        return json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["bf16_tflops"]
    except Exception:  # noqa: BLE001
        return 1590.0


def timed(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    best = 1e9
    for _ in range(3):
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e3 / reps)
    return best


def set_variant(v):
    e = ops.ext()
    e.set_conv_2cta(0); e.set_conv_occ3(1); e.set_persistent_conv(0); e.set_conv_split_producer(False); e.set_conv_tma_store(True)
    nn.USE_HALO3 = False; nn.USE_STRIDED_TMA = False; nn.USE_WGRAD_HALO = True; nn.USE_BN_RECOMPUTE = False
    for tok in v.split("+"):
        if tok == "default":
            pass
        elif tok == "pair":
            e.set_conv_2cta(1)
        elif tok == "notmastore":
            e.set_conv_tma_store(False)
        elif tok == "split":
            e.set_conv_split_producer(True)
        elif tok == "pairdeep":
            e.set_conv_2cta(2)
        elif tok == "recompute":
            nn.USE_BN_RECOMPUTE = True
        elif tok == "occ0":
            e.set_conv_occ3(0)
        elif tok == "occ2":
            e.set_conv_occ3(2)
        elif tok == "persistent":
            e.set_persistent_conv(1)
        elif tok == "halo3":
            nn.USE_HALO3 = True
        elif tok == "strided":
            nn.USE_STRIDED_TMA = True
        elif tok == "nowghalo":
            nn.USE_WGRAD_HALO = False
        else:
            raise SystemExit(f"unknown variant token {tok}")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", default="default,pair")
    ap.add_argument("--layers", default="l1,l2,l3,l4,l2s,l3s,l4s")
    ap.add_argument("--dirs", default="fwd,dgrad,wgrad")
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--eager", action="store_true", help="two plain launches per case, no graph, no timing (for ncu -k captures)")
    a = ap.parse_args()
    pk = peak()
    torch.manual_seed(0)
    rows = []
    for name in a.layers.split(","):
        B, H, Cin, Cout, k, s, p = LAYERS[name]
        Ho = (H + 2 * p - k) // s + 1
        x = torch.randn(B, H, H, Cin, device=DEV).to(BF)
        w = (torch.randn(Cout, k, k, Cin, device=DEV) * 0.05).to(BF)
        y = torch.empty(B, Ho, Ho, Cout, device=DEV, dtype=BF)
        dy = torch.randn(B, Ho, Ho, Cout, device=DEV).to(BF)
        gw = torch.zeros(Cout, k, k, Cin, device=DEV)
        dx = torch.empty_like(x)
        flops = 2.0 * B * Ho * Ho * Cin * Cout * k * k
        for v in a.variants.split(","):
            set_variant(v)
            tag = (name, v)
            fns = {"fwd": lambda: ops.conv2d_fwd_sm100(x, w, None, y, s, p, False, None, tag=tag),
                   "dgrad": lambda: ops.conv2d_dgrad_sm100(dy, w, dx, s, p, False),
                   "wgrad": lambda: ops.conv2d_wgrad_sm100(x, dy, gw, None, s, p, tag=tag, zero=False)}
            ops.conv2d_fwd_sm100(x, w, None, y, s, p, False, None, tag=tag)      # fills the scratch copies wgrad reuses
            for d in a.dirs.split(","):
                try:
                    if a.eager:
                        fns[d](); fns[d](); torch.cuda.synchronize()
                        continue
                    us = timed(fns[d], a.reps)
                    tf = flops / us * 1e-6
                    rows.append((name, d, v, us, tf, tf / pk))
                    print(f"{name:4s} {d:6s} {v:18s} {us:8.1f} us  {tf:7.1f} TFLOP/s  {tf / pk:5.2f} of measured cuBLAS bf16", flush=True)
                except Exception as ex:  # noqa: BLE001
                    print(f"{name:4s} {d:6s} {v:18s} FAILED {type(ex).__name__}: {str(ex)[:100]}", flush=True)
    set_variant("default")
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/bench_convs.json", "w") as fh:
        json.dump([dict(layer=r[0], dir=r[1], variant=r[2], us=r[3], tflops=r[4], frac=r[5]) for r in rows], fh)


if __name__ == "__main__":
    main()
