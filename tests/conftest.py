"""Shared fixtures.  `-m "not gpu"` runs on the CPU build box; `-m gpu` on a B200."""
from __future__ import annotations

import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """Build the product library / oracle if the tree is fresh (no-op when present)."""
    import __graft_entry__ as ge

    ge.ensure_built()


def golden(name: str):
    return np.load(os.path.join(GOLDEN, name))


def rel_to_max(a, b) -> float:
    """max |a-b| / max |b| per last-dim row, maximised over rows -- the parity metric of
    SURVEY.md 8c (element-wise relative error is ill-defined at zero crossings)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    a = a.reshape(-1, a.shape[-1])
    b = b.reshape(-1, b.shape[-1])
    den = np.maximum(np.abs(b).max(axis=-1), 1e-300)
    return float((np.abs(a - b).max(axis=-1) / den).max())
