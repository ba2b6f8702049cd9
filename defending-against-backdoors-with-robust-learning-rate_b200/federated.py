"""Entry point: ``python -m rlr_b200.federated --data=fmnist --num_agents=10 ...`` (reference: ``python federated.py``,
src/federated.py:21-95; same flags).  Multi-GPU: ``torchrun --nproc-per-node 8 -m rlr_b200.federated ...``."""
from __future__ import annotations

from .engine import FLEngine
from .options import args_parser, finalize_args


def main(argv=None):
    args = finalize_args(args_parser(argv))
    engine = FLEngine(args)
    try:
        return engine.fit()
    finally:
        engine.close()


if __name__ == "__main__":
    main()
