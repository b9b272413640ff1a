"""ORACLE: torchsparse/tensor.py (v1.4.0).  Containers only."""
import torch


def _triple(s):
    return tuple(s) if isinstance(s, (tuple, list)) else (s, s, s)


class SparseTensor:
    def __init__(self, feats, coords, stride=1):
        self.feats = feats
        self.coords = coords
        self.stride = _triple(stride)
        self.cmaps = {}
        self.kmaps = {}

    @property
    def F(self):
        return self.feats

    @F.setter
    def F(self, feats):
        self.feats = feats

    @property
    def C(self):
        return self.coords

    @C.setter
    def C(self, coords):
        self.coords = coords

    @property
    def s(self):
        return self.stride

    @s.setter
    def s(self, stride):
        self.stride = _triple(stride)

    def cpu(self):
        self.coords, self.feats = self.coords.cpu(), self.feats.cpu()
        return self

    def cuda(self):
        self.coords, self.feats = self.coords.cuda(), self.feats.cuda()
        return self

    def to(self, device, non_blocking=True):
        self.coords = self.coords.to(device, non_blocking=non_blocking)
        self.feats = self.feats.to(device, non_blocking=non_blocking)
        return self

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

    def cpu(self):
        self.F, self.C = self.F.cpu(), self.C.cpu()
        return self

    def cuda(self):
        self.F, self.C = self.F.cuda(), self.C.cuda()
        return self

    def to(self, device, non_blocking=True):
        self.F = self.F.to(device, non_blocking=non_blocking)
        self.C = self.C.to(device, non_blocking=non_blocking)
        return self

    def __add__(self, other):
        out = PointTensor(self.F + other.F, self.C, self.idx_query, self.weights)
        out.additional_features = self.additional_features
        return out
