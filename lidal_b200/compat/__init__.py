"""Drop-in for the torchsparse==1.4.0 Python surface LiDAL's networks import (boundary B1).

    import sys, lidal_b200.compat as ts
    sys.modules["torchsparse"] = ts; sys.modules["torchsparse.nn"] = ts.nn
    sys.modules["torchsparse.nn.functional"] = ts.nn.functional; sys.modules["torchsparse.nn.utils"] = ts.nn.utils

(`lidal_b200.compat.install()` does exactly that) after which the reference's unmodified
``network/minkunet.py``, ``network/spvcnn.py``, ``network/utils.py``, ``train.py``, ``evaluate.py`` and
``score/prob_inference.py`` run on the hand-written sm_100a kernels.  CUDA tensors only; no CPU fallback.
"""
from .tensor import PointTensor, SparseTensor, cat
from . import nn

__version__ = "1.4.0+lidal_b200"
__all__ = ["SparseTensor", "PointTensor", "cat", "nn", "install"]


def install():
    """Alias this package as ``torchsparse`` for the reference's unmodified imports."""
    import sys
    me = sys.modules[__name__]
    sys.modules["torchsparse"] = me
    sys.modules["torchsparse.nn"] = nn
    sys.modules["torchsparse.nn.functional"] = nn.functional
    sys.modules["torchsparse.nn.utils"] = nn.utils
    return me
