import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(autouse=True)
def _deterministic_pca(monkeypatch):
    """torch.pca_lowrank is randomised; the add_static branch calls it on every forward (MA.py:407).  The suite runs
    with the exact, sign-normalised stand-in the golden add_static case was generated with (tests/util.py)."""
    import torch

    from tests.util import exact_pca_lowrank

    monkeypatch.setattr(torch, "pca_lowrank", exact_pca_lowrank)
