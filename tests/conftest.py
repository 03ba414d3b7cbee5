"""Shared test plumbing.  `-m "not gpu"` tests run in the build container (no GPU); `-m gpu` tests are the
parity tests proper and call the CUDA path through the C-ABI (libbrief_b200.so)."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_gold(name):
    return np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=True)


@pytest.fixture(scope="session")
def gold():
    return load_gold


def packed_params(g, layers, prefix=""):
    """[W0,b0,W1,b1,...] flattened in utils/ModelSave.py order from a golden archive."""
    parts = []
    for l in range(layers):
        parts += [np.asarray(g[f"{prefix}W{l}"]).ravel(), np.asarray(g[f"{prefix}b{l}"]).ravel()]
    return np.concatenate(parts).astype(np.float32)


def unpack(flat, coords_channel, features, layers, out=1):
    """packed -> [(W_l, b_l)]"""
    widths = [coords_channel] + [features] * (layers - 1) + [out]
    res, off = [], 0
    for l in range(layers):
        n = widths[l + 1] * widths[l]
        W = flat[off:off + n].reshape(widths[l + 1], widths[l]); off += n
        b = flat[off:off + widths[l + 1]]; off += widths[l + 1]
        res.append((W, b))
    assert off == flat.size
    return res
