"""Drive the UNMODIFIED reference (baseline/_ref/src) on synthetic data and time its federated round.

What is the reference's and what is ours:
* reference, untouched: ``options.args_parser``, ``utils.distribute_data`` / ``DatasetSplit`` / ``poison_dataset``,
  ``Agent`` (DataLoader, local_train), ``Aggregation.aggregate_updates``, ``models.get_model`` (for its own CNNs).
* shims (ours): ``utils.get_datasets`` -> synthetic torchvision-style dataset (the reference downloads, no network
  here), the 6-line round loop of src/federated.py:65-74 (the reference's is module-level script code that cannot be
  imported), and -- only for ``--model resnet18|resnet34|vgg11|vgg16`` -- ``models.get_model`` -> baseline/torch_models.py because the
  reference has no such architectures.
Timed region = src/federated.py:66-74 (local training of every sampled agent + restore + aggregation), no evaluation.
"""
from __future__ import annotations

import copy
import math
import os
import sys
import time

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.path.join(HERE, "_ref", "src")


def reference_available():
    return os.path.isfile(os.path.join(REF_SRC, "federated.py"))


def _synthetic_vision_dataset(name, n, seed, transform):
    from PIL import Image
    from torch.utils.data import Dataset

    g = torch.Generator().manual_seed(1234567 + seed)
    ncls = 10
    if name == "cifar10":
        h, w, c = 32, 32, 3
    else:
        h, w, c = 28, 28, 1
    coarse = torch.rand(ncls, c, 7, 7, generator=g)
    protos = torch.nn.functional.interpolate(coarse, size=(h, w), mode="bilinear", align_corners=False).permute(0, 2, 3, 1)
    y = (torch.arange(n) % ncls)[torch.randperm(n, generator=g)]
    x = (protos[y] * 0.6 + 0.2 + 0.15 * torch.randn(n, h, w, c, generator=g)).clamp_(0, 1)
    x = (x * 255).round().to(torch.uint8)

    class SynthVision(Dataset):  # same per-sample host path as torchvision's CIFAR10 / FashionMNIST __getitem__
        def __init__(self):
            self.targets = y.clone()                       # LongTensor, like src/utils.py:122
            self.data = x.numpy().copy() if name == "cifar10" else x[..., 0].clone()
            self.transform = transform

        def __len__(self):
            return len(self.targets)

        def __getitem__(self, i):
            img, target = self.data[i], int(self.targets[i])
            img = Image.fromarray(img) if name == "cifar10" else Image.fromarray(img.numpy(), mode="L")
            return self.transform(img), target

    return SynthVision()


def run(data="cifar10", model="resnet18", num_agents=1, local_ep=2, bs=256, aggr="avg", train_size=50000,
        steps=3, warmup=3, theta=0, num_corrupt=0, poison_frac=0.0, device="cuda:0", seed=0, agent_frac=1.0):
    """Returns dict(ms_per_round, rounds_per_s, h2d_bytes_per_round, wall_s)."""
    if not reference_available():
        raise FileNotFoundError("baseline/_ref/src missing; run baseline/install_reference.py")
    cwd = os.getcwd()
    sys.path.insert(0, REF_SRC)
    os.chdir(REF_SRC)  # the reference uses cwd-relative paths (src/utils.py:98,233)
    argv = sys.argv
    try:
        sys.argv = ["federated.py", f"--data={data}", f"--num_agents={num_agents}", f"--local_ep={local_ep}", f"--bs={bs}",
                    f"--aggr={aggr}", f"--robustLR_threshold={theta}", f"--num_corrupt={num_corrupt}",
                    f"--poison_frac={poison_frac}", f"--device={device}", "--snap=1000000", f"--agent_frac={agent_frac}"]
        import utils as ref_utils  # noqa: E402  (reference module)
        import models as ref_models
        from agent import Agent
        from aggregation import Aggregation
        from options import args_parser
        from torchvision import transforms
        from torch.nn.utils import parameters_to_vector, vector_to_parameters

        args = args_parser()
        args.server_lr = args.server_lr if args.aggr == "sign" else 1.0   # src/federated.py:23
        torch.backends.cudnn.enabled = True
        torch.backends.cudnn.benchmark = True                              # src/federated.py:18-19
        torch.manual_seed(seed); np.random.seed(seed)

        if data == "cifar10":
            tf = transforms.Compose([transforms.ToTensor(),
                                     transforms.Normalize(mean=(0.4914, 0.4822, 0.4465), std=(0.2023, 0.1994, 0.2010))])
        else:
            tf = transforms.Compose([transforms.ToTensor(), transforms.Normalize(mean=[0.2860], std=[0.3530])])
        train_dataset = _synthetic_vision_dataset(data, train_size, seed, tf)
        ref_utils.get_datasets = lambda _d: (train_dataset, None)          # shim #1 (documented above)
        if model in ("resnet18", "resnet34", "vgg11", "vgg16"):                                 # shim #3: architectures the reference lacks
            sys.path.insert(0, HERE)
            import torch_models
            ctor = {"resnet18": torch_models.ResNet18, "resnet34": torch_models.ResNet34, "vgg11": torch_models.VGG11,
                    "vgg16": torch_models.VGG16}[model]
            ref_models.get_model = lambda _d: ctor()

        user_groups = ref_utils.distribute_data(train_dataset, args)
        global_model = ref_models.get_model(args.data).to(args.device)
        agents, agent_data_sizes = [], {}
        for _id in range(args.num_agents):
            a = Agent(_id, args, train_dataset, user_groups[_id])
            agent_data_sizes[_id] = a.n_data
            agents.append(a)
        n_params = len(parameters_to_vector(global_model.parameters()))
        aggregator = Aggregation(agent_data_sizes, n_params, None, args, None)
        criterion = torch.nn.CrossEntropyLoss().to(args.device)

        def one_round(rnd):  # src/federated.py:66-74, verbatim control flow
            rnd_global_params = parameters_to_vector(global_model.parameters()).detach()
            agent_updates_dict = {}
            for agent_id in np.random.choice(args.num_agents, math.floor(args.num_agents * args.agent_frac), replace=False):
                update = agents[agent_id].local_train(global_model, criterion)
                agent_updates_dict[agent_id] = update
                vector_to_parameters(copy.deepcopy(rnd_global_params), global_model.parameters())
            aggregator.aggregate_updates(global_model, agent_updates_dict, rnd)

        cuda = str(args.device).startswith("cuda")
        for r in range(warmup):
            one_round(r + 1)
        if cuda:
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
        t0 = time.perf_counter()
        for r in range(steps):
            one_round(warmup + r + 1)
        if cuda:
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
        wall = time.perf_counter() - t0
        if not cuda:
            ms = wall * 1e3 / steps
        px = 3 * 32 * 32 if data == "cifar10" else 28 * 28
        h2d = train_size * local_ep * (px * 4 + 8)  # fp32 batch tensors + int64 labels the reference copies per round
        return {"ms_per_round": ms, "rounds_per_s": 1e3 / ms, "h2d_bytes_per_round": h2d, "wall_s": wall,
                "n_params": n_params}
    finally:
        sys.argv = argv
        os.chdir(cwd)
        for p in (REF_SRC, HERE):
            if p in sys.path:
                sys.path.remove(p)


if __name__ == "__main__":
    import json
    print(json.dumps(run(device="cuda:0" if torch.cuda.is_available() else "cpu", train_size=2000, steps=1, warmup=1)))
