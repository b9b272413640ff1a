"""``torchsparse.nn.utils.get_kernel_offsets`` (used at network/utils.py:69 and by every kernel-map build)."""
import itertools

import torch


def make_ntuple(x, n=3):
    return tuple(x) if isinstance(x, (tuple, list)) else (x,) * n


def get_kernel_offsets(size, stride=1, dilation=1, device="cpu"):
    """int32 [K, 3].  Odd kernel volume: x fastest / z slowest; even volume: x slowest / z fastest."""
    size, stride, dilation = make_ntuple(size), make_ntuple(stride), make_ntuple(dilation)
    axes = [[(i - (size[a] - 1) // 2) * stride[a] * dilation[a] for i in range(size[a])] for a in range(3)]
    if (size[0] * size[1] * size[2]) % 2 == 1:
        offs = [(x, y, z) for z, y, x in itertools.product(axes[2], axes[1], axes[0])]
    else:
        offs = list(itertools.product(axes[0], axes[1], axes[2]))
    return torch.tensor(offs, dtype=torch.int, device=device)
