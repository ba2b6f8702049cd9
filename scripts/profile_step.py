"""Run a few eager (non-graph) local steps of one trainer so ncu can list / profile every kernel of a step.

    ncu --metrics gpu__time_duration.sum --clock-control none -s <skip> -c <count> --csv --log-file out.csv \
        python scripts/profile_step.py --trainer native --steps 4
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from rlr_b200.engine import FLEngine  # noqa: E402
from rlr_b200.options import make_args  # noqa: E402

p = argparse.ArgumentParser()
p.add_argument("--trainer", default="native")
p.add_argument("--model", default="resnet18")
p.add_argument("--data", default="cifar10")
p.add_argument("--steps", type=int, default=4)
p.add_argument("--bs", type=int, default=256)
a = p.parse_args()
args = make_args(data=a.data, model=a.model, num_agents=1, local_ep=1, bs=a.bs, synthetic=a.bs * a.steps, synthetic_val=256, log_dir="",
                 device="cuda:0", trainer=a.trainer, no_graphs=True)
eng = FLEngine(args, verbose=False)
torch.cuda.synchronize()
torch.cuda.nvtx.range_push("round")
eng.run_round(1)
torch.cuda.synchronize()
torch.cuda.nvtx.range_pop()
print("done", eng.trainer.name, eng.round_result())
