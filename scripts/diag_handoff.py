"""Diagnostic: fused hand-off vs round_init path on the FMNIST CNN, per round, for (dropout fused?, agents in flight).
Run: RLR_FUSE_DROPOUT={0,1} python scripts/diag_handoff.py <agents_in_flight>"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from rlr_b200.engine import FLEngine  # noqa: E402
from rlr_b200.options import make_args  # noqa: E402

nf = int(sys.argv[1]) if len(sys.argv) > 1 else 0
model = sys.argv[2] if len(sys.argv) > 2 else "cnn_mnist"


def run(fused):
    args = make_args(data="fmnist" if model != "resnet18" else "cifar10", model=model, num_agents=3, local_ep=1, bs=64, synthetic=192,
                     synthetic_val=64, log_dir="", device="cuda:0", seed=4, no_fused_handoff=not fused, agents_in_flight=nf)
    eng = FLEngine(args, verbose=False)
    snaps = []
    for r in range(1, 4):
        eng.run_round(r)
        snaps.append(eng.global_params().clone())
    torch.cuda.synchronize()
    nt = len(eng.trainers)
    eng.close()
    return snaps, nt


def rr(a, b):
    return float((a.double() - b.double()).pow(2).mean().sqrt() / b.double().pow(2).mean().sqrt())


(a, nt), (b, _), (f, _) = run(False), run(False), run(True)
print(f"fuse_dropout={os.environ.get('RLR_FUSE_DROPOUT', '1')} agents_in_flight={nf} (trainers {nt}) {model}: " +
      " | ".join(f"round {i + 1}: noise {rr(b[i], a[i]):.2e} fused {rr(f[i], a[i]):.2e}" for i in range(3)))
lay = None
d = (f[2] - a[2]).abs()
print("   largest |diff| at", int(d.argmax()), float(d.max()), "n", d.numel(), " nonzero diffs:", int((d > 1e-6).sum()))
