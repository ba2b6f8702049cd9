"""CUDA-event phase timers (the reference only has a tqdm it/s).  Events are recorded on the current
stream; ``elapsed()`` synchronises once at read time, so timing adds no sync inside a round."""
from __future__ import annotations

import time

import torch


class PhaseTimer:
    def __init__(self, device):
        self.cuda = torch.device(device).type == "cuda"
        self.device = device
        self._open = {}
        self._spans = {}

    def start(self, name):
        if self.cuda:
            torch.cuda.nvtx.range_push(name)   # phases also show up as NVTX ranges in ncu / nsys timelines
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self._open[name] = ev
        else:
            self._open[name] = time.perf_counter()

    def stop(self, name):
        if self.cuda:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            torch.cuda.nvtx.range_pop()
            self._spans.setdefault(name, []).append((self._open.pop(name), ev))
        else:
            self._spans.setdefault(name, []).append((self._open.pop(name), time.perf_counter()))

    def elapsed(self, reset=True):
        """{phase: milliseconds} summed over recorded spans."""
        out = {}
        if self.cuda:
            torch.cuda.synchronize(self.device)
        for name, spans in self._spans.items():
            if self.cuda:
                out[name] = sum(a.elapsed_time(b) for a, b in spans)
            else:
                out[name] = sum((b - a) * 1e3 for a, b in spans)
        if reset:
            self._spans = {}
        return out
