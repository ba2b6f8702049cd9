"""End-to-end behaviour on one B200 with the native trainer (qualitative shape of the reference's README plots):
no attack / DBA attack / DBA attack + Robust LR on CIFAR-shaped synthetic data, ResNet-18 and the reference CNN; plus a
native-vs-torch-trainer learning-curve comparison.  Prints one line per round."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rlr_b200.engine import FLEngine
from rlr_b200.options import make_args


def run(tag, rounds=8, **kw):
    base = dict(data="cifar10", model="resnet18", synthetic=8000, synthetic_val=1000, num_agents=8, local_ep=2, bs=128, log_dir="",
                device="cuda:0", seed=3)
    base.update(kw)
    eng = FLEngine(make_args(**base), verbose=False)
    out = []
    for r in range(1, rounds + 1):
        eng.run_round(r)
        ev = eng.evaluate(r)
        loss, flipped = eng.round_result()
        out.append((r, round(ev["val_acc"], 3), round(ev["poison_acc"], 3), round(flipped / eng.layout.n_vote, 3)))
    print(f"{tag:42s} trainer={eng.trainer.name:6s} (round, val_acc, poison_acc, frac_flipped): {out}", flush=True)
    eng.close()


run("resnet18 no attack")
run("resnet18 no attack (torch trainer)", trainer="torch")
run("resnet18 DBA 2/8 corrupt, no defence", num_corrupt=2, poison_frac=0.5)
run("resnet18 DBA 2/8 corrupt, RLR theta=5", num_corrupt=2, poison_frac=0.5, robustLR_threshold=5)
run("cnn_cifar DBA 2/8 corrupt, no defence", model="cnn_cifar", num_corrupt=2, poison_frac=0.5, rounds=12)
run("cnn_cifar DBA 2/8 corrupt, RLR theta=5", model="cnn_cifar", num_corrupt=2, poison_frac=0.5, robustLR_threshold=5, rounds=12)
run("fmnist cnn 1/8 corrupt plus, no defence", data="fmnist", model="cnn_mnist", num_corrupt=1, poison_frac=0.5, rounds=10)
run("fmnist cnn 1/8 corrupt plus, RLR theta=4", data="fmnist", model="cnn_mnist", num_corrupt=1, poison_frac=0.5, robustLR_threshold=4, rounds=10)
