"""ORACLE (test infrastructure): CPU restatement of the torchsparse==1.4.0 Python surface.

Third-party dependency of the reference pinned at docs/requirements.txt:191; its source
is not vendored under /root/reference and it cannot be installed offline, so the
semantics below are restated from its published v1.4.0 behaviour.
PARITY UNPINNED by the reference's own tests (it has none).  Call sites this must satisfy:
network/utils.py:13-102, network/minkunet.py:97-122, network/spvcnn.py:112-155.
"""
from .tensor import PointTensor, SparseTensor
from .operators import cat
from . import nn

__version__ = "1.4.0-oracle"
__all__ = ["SparseTensor", "PointTensor", "cat", "nn"]
