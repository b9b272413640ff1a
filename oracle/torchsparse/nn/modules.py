"""ORACLE: torchsparse/nn/modules/{conv,norm,activation}.py (v1.4.0)."""
import math

import numpy as np
import torch
from torch import nn

from ..tensor import SparseTensor
from . import functional as F
from .utils import make_ntuple


def fapply(x, fn):
    out = SparseTensor(fn(x.feats), x.coords, x.stride)
    out.cmaps, out.kmaps = x.cmaps, x.kmaps
    return out


class Conv3d(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, dilation=1, bias=False, transposed=False):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.kernel_size, self.stride, self.dilation = make_ntuple(kernel_size), make_ntuple(stride), dilation
        self.transposed = transposed
        self.kernel_volume = int(np.prod(self.kernel_size))
        shape = (self.kernel_volume, in_channels, out_channels) if self.kernel_volume > 1 else (in_channels, out_channels)
        self.kernel = nn.Parameter(torch.zeros(*shape))
        self.bias = nn.Parameter(torch.zeros(out_channels)) if bias else None
        self.reset_parameters()

    def reset_parameters(self):
        std = 1.0 / math.sqrt((self.out_channels if self.transposed else self.in_channels) * self.kernel_volume)
        self.kernel.data.uniform_(-std, std)
        if self.bias is not None:
            self.bias.data.uniform_(-std, std)

    def forward(self, x):
        return F.conv3d(x, self.kernel, self.kernel_size, self.bias, self.stride, self.dilation, self.transposed)


class BatchNorm(nn.BatchNorm1d):
    def forward(self, x):
        return fapply(x, super().forward)


class ReLU(nn.ReLU):
    def forward(self, x):
        return fapply(x, super().forward)
