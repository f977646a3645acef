"""pytest configuration: registers the ``gpu`` marker and puts the repo's import roots on sys.path.

``lm-net_b200/`` is a source root (it holds the drop-in ``natten`` package and the
``lmnet_b200`` host package), the repo root makes ``oracle`` importable for the checkers.
"""
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TESTS = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "lm-net_b200"), TESTS):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture
def cpu_backend(monkeypatch):
    """TEST-ONLY: lets the *host-side* Python (modules, autograd plumbing, layout handling) run on a
    machine without a GPU by answering the raw C-ABI launches with the CPU oracle (tests/_cpu_backend.py).
    The product has no such path — outside this fixture CPU tensors raise."""
    import _cpu_backend

    return _cpu_backend.apply(monkeypatch.setattr)
