"""ORACLE: torchsparse/nn/utils/kernel.py (v1.4.0) -- used at network/utils.py:69."""
import numpy as np
import torch


def make_ntuple(x, n=3):
    return tuple(x) if isinstance(x, (tuple, list)) else (x,) * n


def get_kernel_offsets(size, stride=1, dilation=1, device="cpu"):
    size, stride, dilation = make_ntuple(size), make_ntuple(stride), make_ntuple(dilation)
    axes = [np.arange(-size[k] // 2 + 1, size[k] // 2 + 1) * stride[k] * dilation[k] for k in range(3)]
    if np.prod(size) % 2 == 1:      # odd volume: x fastest, z slowest
        offs = [[x, y, z] for z in axes[2] for y in axes[1] for x in axes[0]]
    else:                           # even volume: x slowest, z fastest
        offs = [[x, y, z] for x in axes[0] for y in axes[1] for z in axes[2]]
    return torch.tensor(offs, dtype=torch.int, device=device)
