"""Shared test plumbing.  `-m "not gpu"` runs on the CPU-only build box; `-m gpu` on a B200."""
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
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line("markers", "slow: long-running sweep")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session", autouse=True)
def _built_library():
    """The CUDA library must exist before the package can be imported (no fallback)."""
    import build_native
    build_native.build()


def golden(name):
    with np.load(os.path.join(GOLDEN, name)) as z:
        return {k: torch.from_numpy(z[k]) for k in z.files}


def rel_err(a, b):
    """max |a - b| / max(|b|, rms(b)): elementwise relative error with an rms floor so that
    near-zero bins of a spectrum are judged against the signal level."""
    a, b = a.double(), b.double()
    floor = b.pow(2).mean().sqrt().clamp_min(1e-30)
    return ((a - b).abs() / torch.maximum(b.abs(), floor)).max().item()


def pure_rel_err(a, b):
    a, b = a.double(), b.double()
    return ((a - b).abs() / b.abs().clamp_min(1e-30)).max().item()
