"""Stress of several agents in flight on one GPU: fresh engines (graph capture every time), many rounds, either trainer.
Usage: stress_inflight.py <native|torch> <agents_in_flight> <engines> <rounds>"""
import faulthandler
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from rlr_b200.engine import FLEngine  # noqa: E402
from rlr_b200.options import make_args  # noqa: E402

faulthandler.dump_traceback_later(100, exit=True)       # a deadlock prints every thread's stack and exits
trainer, nf, n_eng, rounds = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
t0 = time.time()
for e in range(n_eng):
    args = make_args(data="fmnist", num_agents=3, local_ep=1, bs=64, synthetic=600, synthetic_val=100, num_corrupt=1, poison_frac=0.5,
                     robustLR_threshold=2, log_dir="", device="cuda:0", trainer=trainer, dtype="fp32" if trainer == "torch" else "bf16",
                     agents_in_flight=nf, seed=e)
    eng = FLEngine(args, verbose=False)
    for r in range(1, rounds + 1):
        eng.run_round(r)
    ev = eng.evaluate(rounds)
    torch.cuda.synchronize()
    eng.close()
    print(f"engine {e}: trainers {len(eng.trainers)} val_acc {ev['val_acc']:.3f}  t={time.time() - t0:.1f}s", flush=True)
print("stress ok", trainer, nf)
