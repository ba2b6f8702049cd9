"""Bisect run-to-run variation of the native trainer: train the same config repeatedly with one op group switched to aten."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rlr_b200.engine import FLEngine
from rlr_b200.options import make_args
from rlr_b200.models import native as nat

model = sys.argv[1] if len(sys.argv) > 1 else "cnn_cifar"
variants = {"all_sm100": {}, "dropout_aten": {"dropout": "aten"}, "only_dropout_sm100": "only_dropout", "all_aten": None}
REPS = 8
orig_init = nat.NativeNet.__init__
for name, over in variants.items():
    def patched(self, layout, device, max_batch, impl="auto", seed=0, act_dtype=None, _over=over):
        orig_init(self, layout, device, max_batch, impl, seed, act_dtype)
        if _over is None or _over == "only_dropout":
            for k in self.impl: self.impl[k] = "aten"
            if _over == "only_dropout": self.impl["dropout"] = "sm100"
        else:
            self.impl.update(_over)
    nat.NativeNet.__init__ = patched
    accs, sums = [], []
    for rep in range(REPS):
        args = make_args(data="cifar10", model=model, num_agents=2, local_ep=2, bs=64, synthetic=1000, synthetic_val=200, log_dir="",
                         device="cuda:0", trainer="native", seed=2)
        eng = FLEngine(args, verbose=False)
        for r in range(1, 4):
            eng.run_round(r)
        accs.append(round(eng.evaluate(3)["val_acc"], 3)); sums.append(round(float(eng.w_global.double().abs().sum()), 4))
        eng.close()
    print(f"{name:14s} acc {accs}  |w| {sums}", flush=True)
