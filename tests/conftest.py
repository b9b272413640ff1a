import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    return {n[:-4]: np.load(os.path.join(GOLDEN, n), allow_pickle=True) for n in os.listdir(GOLDEN) if n.endswith(".npz")}


@pytest.fixture(scope="session")
def oracle_ts():
    """The oracle torchsparse restatement (CPU checker)."""
    import torchsparse
    assert "oracle" in torchsparse.__version__
    return torchsparse


@pytest.fixture(scope="session")
def small_scan():
    """Two NU-shaped views subsampled x4: the nets.npz input."""
    from lidal_b200 import synth
    raw = synth.raycast_scan(42, "NU")
    rs = np.random.RandomState(9)
    return synth.collate_views([synth.score_transform(raw[::4], rs), synth.score_transform(raw[1::4], rs)])
