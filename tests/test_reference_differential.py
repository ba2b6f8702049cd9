"""Differential tests against the reference's own client / evaluation code (src/agent.py, src/utils.py:128-157), on CPU.

The reference Agent is driven with a tiny in-memory dataset whose ``__getitem__`` returns exactly the normalised tensors our
gather kernel's oracle produces; dropout is disabled (p=0) on both sides and the batch is the whole shard, so one local step is
deterministic and the update vectors must agree coordinate by coordinate (after mapping our OHWI / NHWC-flatten layout to the
reference's ``parameters_to_vector`` order)."""
import sys
from types import SimpleNamespace

import pytest
import torch

from rlr_b200.data import make_synthetic
from rlr_b200.models import get_layout, GraphNet
from rlr_b200.options import make_args
from rlr_b200.trainers import TorchTrainer
from rlr_b200.utils import get_loss_n_accuracy


class _RefView(torch.utils.data.Dataset):
    """What the reference expects from a torchvision dataset: ``targets`` + normalised (C,H,W) tensors from __getitem__."""

    def __init__(self, ds):
        self.ds = ds
        self.targets = ds.targets.clone()

    def __len__(self):
        return len(self.ds)

    def __getitem__(self, i):
        x, y = self.ds.batch(torch.tensor([i]))
        return x[0], int(y[0])


@pytest.fixture()
def ref_agent_cls(reference_modules):
    sys.path.insert(0, reference_modules["src"])
    try:
        import importlib
        agent_mod = importlib.import_module("agent")
    finally:
        sys.path.remove(reference_modules["src"])
    return agent_mod.Agent


@pytest.mark.parametrize("data,clip", [("fmnist", 0.0), ("cifar10", 0.0), ("fmnist", 0.02)])
def test_local_train_update_matches_reference_agent(reference_modules, ref_agent_cls, data, clip):
    n = 48
    train, _ = make_synthetic(data, n, 16, seed=4)
    ref_args = SimpleNamespace(num_corrupt=0, bs=n, local_ep=1, client_lr=0.1, client_moment=0.9, clip=clip, device="cpu", num_workers=0,
                               data=data)
    ref_model = reference_modules["models"].get_model(data)
    for m in ref_model.modules():
        if isinstance(m, torch.nn.Dropout2d):
            m.p = 0.0
    ref_agent = ref_agent_cls(0, ref_args, _RefView(train), list(range(n)))
    init_vec = torch.nn.utils.parameters_to_vector(ref_model.parameters()).detach().clone()
    ref_update = ref_agent.local_train(ref_model, torch.nn.CrossEntropyLoss())          # fp64: local - global

    args = make_args(data=data, num_agents=1, local_ep=1, bs=n, clip=clip, log_dir="", device="cpu")
    lay = get_layout(args.model)
    for nd in lay.nodes:
        if nd.op == "dropout":
            nd.attrs["p"] = 0.0
    w_global = torch.zeros(lay.n_total)
    lay.from_reference_vector(init_vec, w_global)
    trainer = TorchTrainer(lay, args, "cpu", n)
    agent = SimpleNamespace(id=0, dataset=train, n_data=n, idxs=torch.arange(n),
                            epoch_indices=lambda seed, rnd, ep: torch.arange(n))
    out = torch.zeros(lay.n_total)
    trainer.train_agent(agent, w_global, out, rnd=1)
    ours = lay.to_reference_vector(out) - lay.to_reference_vector(w_global)
    torch.testing.assert_close(ours.double(), ref_update, atol=2e-6, rtol=1e-4)
    if clip > 0:
        assert float(ours.norm()) <= clip * (1 + 1e-4)


def test_evaluation_matches_reference_function(reference_modules):
    _, val = make_synthetic("cifar10", 64, 90, seed=7)
    lay = get_layout("cnn_cifar")
    w = lay.init_(torch.zeros(lay.n_total), 5)
    net = GraphNet(lay, w, None).eval()
    ref_model = reference_modules["models"].get_model("cifar10").eval()
    torch.nn.utils.vector_to_parameters(lay.to_reference_vector(w), ref_model.parameters())
    loader = torch.utils.data.DataLoader(_RefView(val), batch_size=32, shuffle=False)
    ref_loss, (ref_acc, ref_pc) = reference_modules["utils"].get_loss_n_accuracy(ref_model, torch.nn.CrossEntropyLoss(), loader,
                                                                                 SimpleNamespace(device="cpu"))
    loss, (acc, pc) = get_loss_n_accuracy(lambda x: net(x), val, bs=32)
    assert abs(loss - ref_loss) < 1e-4 and abs(acc - ref_acc) < 1e-9
    torch.testing.assert_close(pc, ref_pc.float(), atol=1e-6, rtol=1e-6, equal_nan=True)
