"""Differential tests against the reference's own client / evaluation code (src/agent.py, src/utils.py:128-157), on CPU.

The reference Agent is driven with a tiny in-memory dataset whose ``__getitem__`` returns exactly the normalised tensors our
gather kernel's oracle produces; dropout is disabled (p=0) on both sides and the batch is the whole shard, so one local step is
deterministic and the update vectors must agree coordinate by coordinate (after mapping our OHWI / NHWC-flatten layout to the
reference's ``parameters_to_vector`` order)."""
import os
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


def _write_fedemnist_tree(ref_utils, root, n_clients=6, per_client=40, seed=0):
    """Files in the layout the reference reads (src/utils.py:105-109, src/agent.py:17), pickled with the REFERENCE's own
    ``utils.H5Dataset`` class from the h5-style ``{client: {'label', 'pixels'}}`` mapping it is built from."""
    import numpy as np
    rs = np.random.RandomState(seed)
    os.makedirs(os.path.join(root, "Fed_EMNIST", "user_trainsets"))
    raw = {f"c{i}": {"label": rs.randint(0, 10, per_client + i), "pixels": rs.rand(per_client + i, 28, 28).astype("float32")}
           for i in range(n_clients + 1)}
    clients = [ref_utils.H5Dataset(raw, f"c{i}") for i in range(n_clients)]
    for i, c in enumerate(clients):
        torch.save(c, os.path.join(root, "Fed_EMNIST", "user_trainsets", f"user_{i}_trainset.pt"))
    allset = ref_utils.H5Dataset(raw, "c0")
    for i in range(1, n_clients):
        allset = allset + ref_utils.H5Dataset(raw, f"c{i}")
    torch.save(allset, os.path.join(root, "Fed_EMNIST", "fed_emnist_all_trainset.pt"))
    torch.save(ref_utils.H5Dataset(raw, f"c{n_clients}"), os.path.join(root, "Fed_EMNIST", "fed_emnist_all_valset.pt"))
    return raw


def test_fedemnist_files_written_by_the_reference_class_load_and_train(reference_modules, tmp_path):
    """Real-data Fed-EMNIST path: per-client ``.pt`` files holding pickled reference ``H5Dataset`` objects are read, the clients
    become index ranges of one device dataset (sizes and pixels preserved), corrupt clients are poisoned, a round trains."""
    import os as _os
    from rlr_b200.engine import FLEngine
    raw = _write_fedemnist_tree(reference_modules["utils"], str(tmp_path))
    args = make_args(data="fedemnist", data_dir=str(tmp_path), num_agents=6, agent_frac=0.5, num_corrupt=1, poison_frac=0.5,
                     pattern_type="plus", local_ep=1, bs=16, rounds=1, log_dir="", device="cpu", seed=3)
    eng = FLEngine(args, verbose=False)
    assert [a.n_data for a in eng.agents] == [40 + i for i in range(6)] and eng.n_part == 3
    assert len(eng.val_dataset) == 46 and eng.val_dataset.data.dtype == torch.float32
    a3 = eng.agents[3]
    got = eng.train_dataset.data[a3.idxs].reshape(-1, 28, 28)
    assert torch.equal(got, torch.as_tensor(raw["c3"]["pixels"]))             # honest client: pixels untouched, NHWC view
    assert torch.equal(eng.train_dataset.targets[a3.idxs], torch.as_tensor(raw["c3"]["label"]).long())
    n_base = int((torch.as_tensor(raw["c0"]["label"]) == args.base_class).sum())
    assert len(eng.agents[0].poisoned_idxs) == n_base // 2                    # floor(poison_frac * |base class in shard|)
    info = eng.run_round(1)
    assert len(info["chosen"]) == 3
    ev = eng.evaluate(1)
    assert ev["val_loss"] == ev["val_loss"]
    eng.close()


def test_fedemnist_cli_without_the_reference_on_the_path(reference_modules, tmp_path):
    """Same files through ``python federated.py`` in a fresh interpreter, where pickle's ``utils.H5Dataset`` must resolve to our
    shim class (the reference's modules are not importable there)."""
    import subprocess
    _write_fedemnist_tree(reference_modules["utils"], str(tmp_path))
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "federated.py", "--data=fedemnist", f"--data_dir={tmp_path}", "--num_agents=6", "--agent_frac=0.5",
                        "--num_corrupt=1", "--poison_frac=0.5", "--local_ep=1", "--bs=16", "--rounds=2", "--snap=1", "--device=cpu",
                        "--log_dir=", "--no_tensorboard"], cwd=root, env=dict(os.environ, CUDA_VISIBLE_DEVICES=""),
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    assert "| Val_Loss/Val_Acc:" in r.stdout and "Training has finished!" in r.stdout
