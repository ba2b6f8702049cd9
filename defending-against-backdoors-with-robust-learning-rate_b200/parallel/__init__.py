"""Multi-GPU layer: one process per GPU (agent k <-> GPU k), torch.distributed for rendezvous/plumbing, and
hand-written NVLink kernels for the data path.  The reference has no counterpart: its
"communication" is a Python dict on one device (src/federated.py:67-74)."""
from .comm import DistContext, init_distributed
from .symm import SymmetricBuffer
from .fused_agg import FusedAggregator

__all__ = ["DistContext", "init_distributed", "SymmetricBuffer", "FusedAggregator"]
