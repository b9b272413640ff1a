"""``spnn.Conv3d / BatchNorm / ReLU`` with torchsparse-1.4.0 parameter names and shapes so the reference's
checkpoints load with ``strict=True`` (train.py:66-68, score/prob_inference.py:66-71)."""
from __future__ import annotations

import math

import torch
from torch import nn

from ..tensor import SparseTensor
from . import functional as F
from .utils import make_ntuple

__all__ = ["Conv3d", "BatchNorm", "ReLU"]


def fapply(x: SparseTensor, fn) -> SparseTensor:
    out = SparseTensor(fn(x.feats), x.coords, x.stride)
    out.cmaps, out.kmaps = x.cmaps, x.kmaps
    return out


class Conv3d(nn.Module):
    """Parameter ``kernel``: [K, Cin, Cout], or [Cin, Cout] when the kernel volume is 1; no bias by default."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, dilation=1, bias=False, transposed=False):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.dilation = make_ntuple(kernel_size), make_ntuple(stride), dilation
        self.transposed = transposed
        self.kernel_volume = self.kernel_size[0] * self.kernel_size[1] * self.kernel_size[2]
        shape = (in_channels, out_channels) if self.kernel_volume == 1 else (self.kernel_volume, in_channels, out_channels)
        self.kernel = nn.Parameter(torch.zeros(shape))
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        fan = (self.out_channels if self.transposed else self.in_channels) * self.kernel_volume
        bound = 1.0 / math.sqrt(fan)
        with torch.no_grad():
            self.kernel.uniform_(-bound, bound)
            if self.bias is not None:
                self.bias.uniform_(-bound, bound)

    def extra_repr(self):
        return (f"{self.in_channels}, {self.out_channels}, kernel_size={self.kernel_size}, stride={self.stride}"
                + (", transposed=True" if self.transposed else ""))

    def forward(self, x: SparseTensor) -> SparseTensor:
        return F.conv3d(x, self.kernel, self.kernel_size, self.bias, self.stride, self.dilation, self.transposed)


class BatchNorm(nn.BatchNorm1d):
    """isinstance(nn.BatchNorm1d) is relied on by network/minkunet.py:93."""

    def forward(self, x: SparseTensor) -> SparseTensor:
        return fapply(x, super().forward)


class ReLU(nn.ReLU):
    def forward(self, x: SparseTensor) -> SparseTensor:
        return fapply(x, super().forward)
