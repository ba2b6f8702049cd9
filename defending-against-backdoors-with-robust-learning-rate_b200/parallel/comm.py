"""Process-group bootstrap.  ``torchrun`` (or any launcher that sets RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*)
starts one process per GPU; NCCL is used on GPUs, gloo on CPU (the plumbing config of BASELINE.json).

The reference has no counterpart: it is one process on one device (``--device``, src/options.py:67-68) that trains its agents one
after another (src/federated.py:68-72)."""
from __future__ import annotations

import os
from dataclasses import dataclass

import torch
import torch.distributed as dist


@dataclass
class DistContext:
    rank: int = 0
    world: int = 1
    local_rank: int = 0
    device: torch.device = torch.device("cpu")
    backend: str = "none"
    single_node: bool = True      # every rank on one host: peer-mapped symmetric memory (NVLink / NVSwitch) is available

    @property
    def is_dist(self):
        return self.world > 1

    @property
    def is_main(self):
        return self.rank == 0

    def barrier(self):
        if self.is_dist:
            if self.backend == "nccl":
                dist.barrier(device_ids=[self.device.index])
            else:
                dist.barrier()

    def all_reduce_sum(self, t):
        if self.is_dist:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return t

    def all_reduce_max(self, t):
        if self.is_dist:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t

    def all_gather(self, t):
        """[world, *t.shape] tensor of every rank's ``t``."""
        if not self.is_dist:
            return t.unsqueeze(0)
        flat = t.contiguous().view(-1)
        out = torch.empty(self.world * flat.numel(), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(out, flat)
        return out.view(self.world, *t.shape)

    def all_gather_object(self, obj):
        if not self.is_dist:
            return [obj]
        out = [None] * self.world
        dist.all_gather_object(out, obj)
        return out

    def broadcast(self, t, src=0):
        if self.is_dist:
            dist.broadcast(t, src)
        return t


def init_distributed(device=None, backend: str | None = None) -> DistContext:
    """Create the context from the launcher's environment; single-process if WORLD_SIZE is unset/1."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", str(rank)))
    use_cuda = torch.cuda.is_available() and (device is None or str(device).startswith("cuda"))
    if world > 1:
        dev = torch.device(f"cuda:{local_rank}") if use_cuda else torch.device("cpu")
    else:
        dev = torch.device(device) if device is not None else torch.device("cuda:0" if use_cuda else "cpu")
    if dev.type == "cuda":
        torch.cuda.set_device(dev)
    if world == 1:
        return DistContext(0, 1, 0, dev, "none")
    be = backend if backend in ("nccl", "gloo") else ("nccl" if dev.type == "cuda" else "gloo")
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29511")
    if not dist.is_initialized():
        kw = {"device_id": dev} if be == "nccl" else {}
        dist.init_process_group(backend=be, rank=rank, world_size=world, **kw)
    # The fused aggregation kernel addresses every peer's memory directly, which needs all ranks under one NVSwitch domain;
    # jobs that span hosts fall back to the NCCL all-gather transport (FusedAggregator, backend "auto").
    local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    return DistContext(rank, world, local_rank, dev, be, single_node=(local_world >= world))
