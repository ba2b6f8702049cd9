"""Models: the reference's CNNs (src/models.py) plus ResNet-18 / VGG-11, as IR programs over flat buffers."""
from __future__ import annotations

import torch

from .graph import FlatLayout, GraphNet, Node, ParamInfo
from .zoo import ZOO

DATA_TO_MODEL = {"fmnist": "cnn_mnist", "fedemnist": "cnn_mnist", "cifar10": "cnn_cifar"}


def get_layout(model: str) -> FlatLayout:
    """Flat layout of a named architecture (``model`` in ZOO) or of a dataset's default model
    (reference ``get_model(data)`` mapping, src/models.py:4-8)."""
    name = DATA_TO_MODEL.get(model, model)
    if name not in ZOO:
        raise ValueError(f"unknown model {model!r}; choose from {sorted(ZOO)}")
    nodes, in_shape = ZOO[name]()
    lay = FlatLayout(nodes, in_shape)
    lay.name = name
    return lay


def get_model(data_or_name: str, device="cpu", seed=0, compute_dtype=torch.float32) -> GraphNet:
    """A ready-to-train torch module (reference ``get_model``, src/models.py:4) with freshly initialised flat
    parameter and gradient buffers on ``device``."""
    lay = get_layout(data_or_name)
    w = torch.zeros(lay.n_total, dtype=torch.float32, device=device)
    g = torch.zeros(lay.n_total, dtype=torch.float32, device=device)
    lay.init_(w, seed)
    return GraphNet(lay, w, g, compute_dtype)


__all__ = ["FlatLayout", "GraphNet", "Node", "ParamInfo", "ZOO", "get_layout", "get_model"]
