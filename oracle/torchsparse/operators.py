"""ORACLE: torchsparse/operators.py (v1.4.0) -- network/minkunet.py:106-118."""
import torch
from .tensor import SparseTensor


def cat(inputs):
    out = SparseTensor(torch.cat([x.feats for x in inputs], dim=1), inputs[0].coords, inputs[0].stride)
    out.cmaps, out.kmaps = inputs[0].cmaps, inputs[0].kmaps
    return out
