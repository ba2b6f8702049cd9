"""Checkpoint / resume (absent in the reference): flat global parameters + round + RNG state."""
from __future__ import annotations

import os
import random

import numpy as np
import torch


def save_checkpoint(path, w_global, rnd, args, layout, extra=None):
    tmp = path + ".tmp"
    os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
    torch.save({
        "w_global": w_global.detach().cpu(),
        "round": int(rnd),
        "args": {k: (str(v) if isinstance(v, torch.device) else v) for k, v in vars(args).items()},
        "model": getattr(layout, "name", ""),
        "n_params": layout.n_params, "n_vote": layout.n_vote, "n_total": layout.n_total,
        "rng": {"torch": torch.get_rng_state(), "numpy": np.random.get_state(), "python": random.getstate()},
        "extra": extra or {},
    }, tmp)
    os.replace(tmp, path)


def load_checkpoint(path, w_global, layout, restore_rng=True):
    ck = torch.load(path, map_location="cpu", weights_only=False)
    if ck["n_total"] != layout.n_total or ck["n_params"] != layout.n_params:
        raise ValueError(f"checkpoint is for a different model ({ck.get('model')!r}: {ck['n_params']} params)")
    w_global.copy_(ck["w_global"].to(w_global.device))
    if restore_rng:
        torch.set_rng_state(ck["rng"]["torch"])
        np.random.set_state(ck["rng"]["numpy"])
        random.setstate(ck["rng"]["python"])
    return ck
