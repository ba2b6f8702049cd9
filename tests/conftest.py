import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
REF_SRC = "/root/reference/src"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        # a GPU test that deadlocks (spin-wait protocols, graph capture) must fail with a stack dump, not stall the whole suite:
        # pytest-timeout's "thread" method also fires while the main thread sits inside a CUDA synchronise
        if config.pluginmanager.hasplugin("timeout"):
            for it in items:
                if "gpu" in it.keywords and it.get_closest_marker("timeout") is None:
                    it.add_marker(pytest.mark.timeout(300, method="thread"))
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def reference_modules():
    """The reference's own modules (utils / aggregation / agent) for differential tests; skipped if not mounted.
    They are imported under their bare names with cwd=src semantics, like the reference requires."""
    src = REF_SRC if os.path.isdir(REF_SRC) else os.path.join(ROOT, "baseline", "_ref", "src")
    if not os.path.isdir(src):
        pytest.skip("reference sources not available")
    sys.path.insert(0, src)
    try:
        import importlib
        mods = {name: importlib.import_module(name) for name in ("utils", "aggregation", "models")}
    finally:
        sys.path.remove(src)
    mods["src"] = src
    return mods
