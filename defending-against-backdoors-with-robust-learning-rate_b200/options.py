"""Command-line interface.

Same flag names, types and defaults as the reference (src/options.py:4-74) so existing command lines
(src/runner.sh:12-38) keep working, plus engine flags that have no reference counterpart
(``--model --dtype --backend --synthetic --seed --checkpoint ...``).  ``finalize_args`` applies the
reference's post-parse fix-up ``server_lr = server_lr if aggr == 'sign' else 1.0`` (src/federated.py:23).
"""
from __future__ import annotations

import argparse

import torch

DATASETS = ("fmnist", "fedemnist", "cifar10")
AGGREGATORS = ("avg", "comed", "sign")
PATTERNS = ("plus", "square", "copyright", "apple")
MODELS = ("auto", "cnn_mnist", "cnn_cifar", "resnet18", "resnet34", "vgg11", "vgg16")


def _default_device():
    return "cuda:0" if torch.cuda.is_available() else "cpu"


def build_parser() -> argparse.ArgumentParser:
    p = argparse.ArgumentParser(
        description="Federated learning with the Robust Learning Rate backdoor defense (B200-native engine)")
    # ---- reference flags (names/defaults: src/options.py:7-70) ----
    p.add_argument("--data", type=str, default="fmnist", help="dataset: fmnist | fedemnist | cifar10")
    p.add_argument("--num_agents", type=int, default=10, help="number of agents: K")
    p.add_argument("--agent_frac", type=float, default=1, help="fraction of agents per round: C")
    p.add_argument("--num_corrupt", type=int, default=0, help="number of corrupt agents (ids 0..n-1)")
    p.add_argument("--rounds", type=int, default=200, help="number of communication rounds: R")
    p.add_argument("--aggr", type=str, default="avg", help="aggregation rule: avg | comed | sign")
    p.add_argument("--local_ep", type=int, default=2, help="number of local epochs: E")
    p.add_argument("--bs", type=int, default=256, help="local batch size: B")
    p.add_argument("--client_lr", type=float, default=0.1, help="clients' learning rate")
    p.add_argument("--client_moment", type=float, default=0.9, help="clients' momentum")
    p.add_argument("--server_lr", type=float, default=1, help="server learning rate (only honoured for aggr=sign)")
    p.add_argument("--base_class", type=int, default=5, help="base class of the backdoor attack")
    p.add_argument("--target_class", type=int, default=7, help="target class of the backdoor attack")
    p.add_argument("--poison_frac", type=float, default=0.0, help="fraction of base-class samples a corrupt agent poisons")
    p.add_argument("--pattern_type", type=str, default="plus", help="trojan pattern: plus | square | copyright | apple")
    p.add_argument("--robustLR_threshold", type=int, default=0, help="RLR sign-vote threshold theta (0 = defense off)")
    p.add_argument("--clip", type=float, default=0, help="L2 ball radius for client PGD / noise scale (0 = off)")
    p.add_argument("--noise", type=float, default=0, help="server noise multiplier: std = noise*clip (0 = off)")
    p.add_argument("--top_frac", type=int, default=100, help="top-k Fisher coordinates in the sign-agreement diagnostic")
    p.add_argument("--snap", type=int, default=1, help="evaluate every `snap` rounds")
    p.add_argument("--device", default=_default_device(), help="device (single-process mode); ranks use cuda:LOCAL_RANK")
    p.add_argument("--num_workers", type=int, default=0, help="accepted for CLI compatibility; data is device-resident")
    # ---- engine flags (no reference counterpart) ----
    p.add_argument("--model", type=str, default="auto", choices=MODELS,
                   help="auto = reference mapping (fmnist/fedemnist->cnn_mnist, cifar10->cnn_cifar)")
    p.add_argument("--dtype", type=str, default="bf16", choices=("fp32", "bf16"),
                   help="activation/GEMM-operand dtype on GPU (master params, updates, aggregation stay fp32)")
    p.add_argument("--backend", type=str, default="auto", choices=("auto", "fused", "nccl", "gloo", "local"),
                   help="aggregation transport: fused = P2P/multicast sm_100a kernel; nccl/gloo = all_gather + kernel")
    p.add_argument("--agents_in_flight", type=int, default=0,
                   help="agents a GPU trains CONCURRENTLY when it hosts several per round (one trainer + CUDA stream each); helps the "
                        "small launch-bound CNNs, costs one set of activation buffers per extra agent.  0 = auto (2 for models under 4 M "
                        "parameters on a GPU, else 1)")
    p.add_argument("--no_fused_handoff", action="store_true",
                   help="keep the separate round_init pass and the aggregation kernel's barrier-out instead of fusing the parameter "
                        "hand-off of a round with the first local GEMM (native trainer; on by default)")
    p.add_argument("--agg_transport", type=str, default="auto", choices=("auto", "gather", "reduce"),
                   help="nccl/gloo back-ends only: gather = all_gather every participant's parameters (needed for comed); reduce = "
                        "all_reduce per-coordinate vote / weighted-sum partials (avg, sign, RLR: O(N) traffic per rank); "
                        "auto = reduce when the ranks span several hosts")
    p.add_argument("--trainer", type=str, default="auto", choices=("auto", "native", "torch"),
                   help="local-training executor: native = sm_100a kernels, torch = autograd oracle (CPU / baseline)")
    p.add_argument("--class_per_agent", type=int, default=10,
                   help="classes per agent in the partitioner (10 = IID like the reference's calls; fewer = label-skewed non-IID; "
                        "reference distribute_data parameter, src/utils.py:58)")
    p.add_argument("--synthetic", type=int, default=0, help=">0: use a synthetic dataset with this many training samples")
    p.add_argument("--synthetic_val", type=int, default=0, help="synthetic validation-set size (default train/5)")
    p.add_argument("--data_dir", type=str, default="../data", help="dataset root (reference: '../data', src/utils.py:98)")
    p.add_argument("--seed", type=int, default=0, help="seed for init / sampling / shuffling / dropout")
    p.add_argument("--log_dir", type=str, default="logs", help="TensorBoard/JSONL root (reference: 'logs/')")
    p.add_argument("--no_tensorboard", action="store_true", help="JSONL + stdout only")
    p.add_argument("--checkpoint", type=str, default="", help="checkpoint path (written every --ckpt_every rounds)")
    p.add_argument("--ckpt_every", type=int, default=0, help="0 = only at the end (if --checkpoint is set)")
    p.add_argument("--resume", type=str, default="", help="resume from this checkpoint")
    p.add_argument("--diagnostics", action="store_true",
                   help="enable the reference's disabled diagnostics (update norms, Fisher sign agreement)")
    p.add_argument("--server_clip", action="store_true",
                   help="clip each update to L2 norm --clip on the server (reference clip_updates, dead code there)")
    p.add_argument("--no_graphs", action="store_true", help="do not capture the local step in CUDA graphs")
    p.add_argument("--profile_phases", action="store_true", help="print per-phase CUDA-event timings every round")
    return p


def finalize_args(args: argparse.Namespace) -> argparse.Namespace:
    """Post-parse normalisation shared by the CLI and programmatic users."""
    # reference src/federated.py:23 -- server_lr is forced to 1 unless sign aggregation is used
    args.server_lr = args.server_lr if args.aggr == "sign" else 1.0
    if args.model == "auto":
        args.model = "cnn_cifar" if args.data == "cifar10" else "cnn_mnist"
    if args.aggr not in AGGREGATORS:
        # the reference silently aggregates to 0 (src/aggregation.py:26); we refuse instead
        raise ValueError(f"unknown --aggr {args.aggr!r}; expected one of {AGGREGATORS}")
    if args.data not in DATASETS:
        raise ValueError(f"unknown --data {args.data!r}; expected one of {DATASETS}")
    return args


def args_parser(argv=None) -> argparse.Namespace:
    """Parse flags the way the reference's ``args_parser`` does (src/options.py:4)."""
    return build_parser().parse_args(argv)


def make_args(**overrides) -> argparse.Namespace:
    """Programmatic construction: defaults + overrides, already finalised."""
    args = build_parser().parse_args([])
    for k, v in overrides.items():
        if not hasattr(args, k):
            raise AttributeError(f"unknown option {k!r}")
        setattr(args, k, v)
    return finalize_args(args)


def print_exp_details(args) -> None:
    """Experiment banner; same 14 fields as reference src/utils.py:287-303 plus engine fields."""
    print("======================================")
    print(f"    Dataset: {args.data}")
    print(f"    Global Rounds: {args.rounds}")
    print(f"    Aggregation Function: {args.aggr}")
    print(f"    Number of agents: {args.num_agents}")
    print(f"    Fraction of agents: {args.agent_frac}")
    print(f"    Batch size: {args.bs}")
    print(f"    Client_LR: {args.client_lr}")
    print(f"    Server_LR: {args.server_lr}")
    print(f"    Client_Momentum: {args.client_moment}")
    print(f"    RobustLR_threshold: {args.robustLR_threshold}")
    print(f"    Noise Ratio: {args.noise}")
    print(f"    Number of corrupt agents: {args.num_corrupt}")
    print(f"    Poison Frac: {args.poison_frac}")
    print(f"    Clip: {args.clip}")
    print(f"    Model / dtype / trainer / backend: {args.model} / {args.dtype} / {args.trainer} / {args.backend}")
    print("======================================")
