"""Real-data path of ``get_datasets`` (reference ``utils.get_datasets``, src/utils.py:95-124): files in torchvision's on-disk layout
are written into a temporary ``data_dir`` and read back through the engine's loader -- no network, no synthetic substitution --
then a whole CLI run trains on them.  FashionMNIST files are also read with torchvision's own ``read_image_file`` to pin the format."""
import gzip
import os
import pickle
import struct
import subprocess
import sys

import numpy as np
import pytest
import torch

import rlr_b200  # noqa: F401
from rlr_b200.data import get_datasets

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _write_idx(path, arr, gz=False):
    arr = np.ascontiguousarray(arr, dtype=np.uint8)
    blob = struct.pack(">HBB", 0, 0x08, arr.ndim) + struct.pack(">" + "I" * arr.ndim, *arr.shape) + arr.tobytes()
    with (gzip.open(path + ".gz", "wb") if gz else open(path, "wb")) as f:
        f.write(blob)


def _make_fmnist(root, n_train=300, n_val=80, gz=False):
    raw = os.path.join(root, "FashionMNIST", "raw")
    os.makedirs(raw, exist_ok=True)
    rs = np.random.RandomState(0)
    out = {}
    for split, n in (("train", n_train), ("t10k", n_val)):
        x = rs.randint(0, 256, (n, 28, 28), dtype=np.uint8)
        y = (np.arange(n) % 10).astype(np.uint8)
        _write_idx(os.path.join(raw, f"{split}-images-idx3-ubyte"), x, gz)
        _write_idx(os.path.join(raw, f"{split}-labels-idx1-ubyte"), y, gz)
        out[split] = (x, y)
    return out


def _make_cifar(root, per_batch=40, n_val=50):
    base = os.path.join(root, "cifar-10-batches-py")
    os.makedirs(base, exist_ok=True)
    rs = np.random.RandomState(1)
    out = {"train": [], "test": None}
    for i in range(1, 6):
        d = {"data": rs.randint(0, 256, (per_batch, 3072), dtype=np.uint8), "labels": [int(v) for v in rs.randint(0, 10, per_batch)]}
        with open(os.path.join(base, f"data_batch_{i}"), "wb") as f:
            pickle.dump(d, f)
        out["train"].append(d)
    d = {"data": rs.randint(0, 256, (n_val, 3072), dtype=np.uint8), "labels": [int(v) for v in rs.randint(0, 10, n_val)]}
    with open(os.path.join(base, "test_batch"), "wb") as f:
        pickle.dump(d, f)
    out["test"] = d
    return out


@pytest.mark.parametrize("gz", [False, True])
def test_fmnist_files_in_torchvision_layout(tmp_path, gz):
    ref = _make_fmnist(str(tmp_path), gz=gz)
    tr, va = get_datasets("fmnist", str(tmp_path))
    assert tr.data.shape == (300, 28, 28, 1) and tr.data.dtype == torch.uint8 and va.data.shape == (80, 28, 28, 1)
    assert torch.equal(tr.data[..., 0], torch.from_numpy(ref["train"][0])) and torch.equal(tr.targets, torch.from_numpy(ref["train"][1]).long())
    assert torch.equal(va.data[..., 0], torch.from_numpy(ref["t10k"][0]))
    if not gz:   # the same bytes through torchvision's reader of that layout
        mnist = pytest.importorskip("torchvision.datasets.mnist")
        tv = mnist.read_image_file(os.path.join(str(tmp_path), "FashionMNIST", "raw", "train-images-idx3-ubyte"))
        assert torch.equal(tv, tr.data[..., 0])


def test_cifar10_files_in_torchvision_layout(tmp_path):
    ref = _make_cifar(str(tmp_path))
    tr, va = get_datasets("cifar10", str(tmp_path))
    assert tr.data.shape == (200, 32, 32, 3) and tr.targets.dtype == torch.int64 and len(va) == 50
    want = np.concatenate([d["data"] for d in ref["train"]]).reshape(-1, 3, 32, 32).transpose(0, 2, 3, 1)   # torchvision: HWC uint8
    assert torch.equal(tr.data, torch.from_numpy(np.ascontiguousarray(want)))
    assert tr.targets.tolist() == [v for d in ref["train"] for v in d["labels"]]


def test_missing_files_raise_instead_of_substituting_synthetic_data(tmp_path):
    with pytest.raises(FileNotFoundError, match="--synthetic"):
        get_datasets("cifar10", str(tmp_path))
    with pytest.raises(FileNotFoundError):
        get_datasets("fmnist", str(tmp_path))


def test_cli_trains_on_real_files(tmp_path):
    """`python federated.py --data fmnist --data_dir <dir>` end to end on on-disk files (no --synthetic)."""
    _make_fmnist(str(tmp_path), n_train=400, n_val=100)
    cmd = [sys.executable, os.path.join(ROOT, "federated.py"), "--data=fmnist", f"--data_dir={tmp_path}", "--num_agents=2", "--rounds=1",
           "--local_ep=1", "--bs=64", "--device=cpu", f"--log_dir={tmp_path}/logs", "--no_tensorboard"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    assert "Val_Loss/Val_Acc" in out.stdout and "synthetic" not in out.stdout.lower()
