"""SparseTensor / PointTensor / cat with the attribute surface the reference touches
(network/utils.py:27-31,57-59,84-92,171; network/spvcnn.py:114,120,131; network/minkunet.py:106-118)."""
from __future__ import annotations

import torch


def _triple(s):
    return tuple(s) if isinstance(s, (tuple, list)) else (s, s, s)


class SparseTensor:
    """feats float [N, C]; coords int32 [N, 4] = (x, y, z, batch); stride 3-tuple; shared cmaps / kmaps dicts."""

    def __init__(self, feats, coords, stride=1):
        self.feats = feats
        self.coords = coords
        self.stride = _triple(stride)
        self.cmaps = {}
        self.kmaps = {}

    F = property(lambda self: self.feats, lambda self, v: setattr(self, "feats", v))
    C = property(lambda self: self.coords, lambda self, v: setattr(self, "coords", v))
    s = property(lambda self: self.stride, lambda self, v: setattr(self, "stride", _triple(v)))

    def _moved(self, fn):
        self.coords, self.feats = fn(self.coords), fn(self.feats)
        return self

    def cpu(self):
        return self._moved(lambda t: t.cpu())

    def cuda(self):
        return self._moved(lambda t: t.cuda())

    def to(self, device, non_blocking=True):
        return self._moved(lambda t: t.to(device, non_blocking=non_blocking))

    def __add__(self, other):
        out = SparseTensor(self.feats + other.feats, self.coords, self.stride)
        out.cmaps, out.kmaps = self.cmaps, self.kmaps
        return out


class PointTensor:
    def __init__(self, feats, coords, idx_query=None, weights=None):
        self.F = feats
        self.C = coords
        self.idx_query = idx_query if idx_query is not None else {}
        self.weights = weights if weights is not None else {}
        self.additional_features = {"idx_query": {}, "counts": {}}

    def _moved(self, fn):
        self.F, self.C = fn(self.F), fn(self.C)
        return self

    def cpu(self):
        return self._moved(lambda t: t.cpu())

    def cuda(self):
        return self._moved(lambda t: t.cuda())

    def to(self, device, non_blocking=True):
        return self._moved(lambda t: t.to(device, non_blocking=non_blocking))

    def __add__(self, other):
        out = PointTensor(self.F + other.F, self.C, self.idx_query, self.weights)
        out.additional_features = self.additional_features
        return out


def cat(inputs):
    out = SparseTensor(torch.cat([t.feats for t in inputs], dim=1), inputs[0].coords, inputs[0].stride)
    out.cmaps, out.kmaps = inputs[0].cmaps, inputs[0].kmaps
    return out
