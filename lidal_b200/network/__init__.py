"""Host-side mirrors of the reference's two networks, parametrised by a sparse backend.

The reference's ``network/minkunet.py:14-122`` and ``network/spvcnn.py:9-155`` are *callers* of
the drop-in boundary and stay unmodified in a real deployment (alias ``lidal_b200.compat`` as
``torchsparse``).  The GPU box has no ``/root/reference``, so tests and ``bench.py`` need
definitions living in this repo: the same U-Net is described once as a table
(``UNET_CHANNELS``) and instantiated against any backend exposing the torchsparse-1.4.0 surface
(``lidal_b200.compat`` on CUDA, ``oracle/torchsparse`` on CPU).  Module attribute names follow the
reference so ``state_dict`` keys/shapes are identical (checked against the reference's own
classes by ``tests/test_oracle.py`` and ``tests/test_abi.py`` when ``/root/reference`` is present).
"""
from __future__ import annotations

import hashlib

import torch
from torch import nn

from .point_voxel import PointVoxel

# cs = [32, 32, 64, 128, 256, 256, 128, 96, 96] at cr = 1.0   (network/minkunet.py:18-20)
UNET_CHANNELS = (32, 32, 64, 128, 256, 256, 128, 96, 96)
IN_CHANNELS = 4


def _conv_bn_relu(be, cin, cout, ks, stride, transposed=False):
    kw = dict(kernel_size=ks, stride=stride)
    if transposed:
        kw["transposed"] = True
    else:
        kw["dilation"] = 1
    return nn.Sequential(be.nn.Conv3d(cin, cout, **kw), be.nn.BatchNorm(cout), be.nn.ReLU(True))


class _Down(nn.Module):
    """k2 s2 strided conv block (network/utils.py:105-121)."""

    def __init__(self, be, c):
        super().__init__()
        self.net = _conv_bn_relu(be, c, c, 2, 2)

    def forward(self, x):
        return self.net(x)


class _Up(nn.Module):
    """k2 s2 transposed conv block (network/utils.py:124-139)."""

    def __init__(self, be, cin, cout):
        super().__init__()
        self.net = _conv_bn_relu(be, cin, cout, 2, 2, transposed=True)

    def forward(self, x):
        return self.net(x)


class _Res(nn.Module):
    """Two k3 convs + identity / 1x1 skip (network/utils.py:142-172)."""

    def __init__(self, be, cin, cout):
        super().__init__()
        self.net = nn.Sequential(
            be.nn.Conv3d(cin, cout, kernel_size=3, dilation=1, stride=1), be.nn.BatchNorm(cout), be.nn.ReLU(True),
            be.nn.Conv3d(cout, cout, kernel_size=3, dilation=1, stride=1), be.nn.BatchNorm(cout))
        self.downsample = nn.Identity() if cin == cout else nn.Sequential(
            be.nn.Conv3d(cin, cout, kernel_size=1, dilation=1, stride=1), be.nn.BatchNorm(cout))
        self.relu = be.nn.ReLU(True)

    def forward(self, x):
        return self.relu(self.net(x) + self.downsample(x))


class _UNetTrunk(nn.Module):
    def __init__(self, be, class_num):
        super().__init__()
        self.be = [be]                      # list: keep the backend module out of nn.Module registration
        cs = UNET_CHANNELS
        self.stem = nn.Sequential(
            be.nn.Conv3d(IN_CHANNELS, cs[0], kernel_size=3, stride=1), be.nn.BatchNorm(cs[0]), be.nn.ReLU(True),
            be.nn.Conv3d(cs[0], cs[0], kernel_size=3, stride=1), be.nn.BatchNorm(cs[0]), be.nn.ReLU(True))
        for i in range(4):                  # encoder: stage1..4
            setattr(self, f"stage{i + 1}", nn.Sequential(
                _Down(be, cs[i]), _Res(be, cs[i], cs[i + 1]), _Res(be, cs[i + 1], cs[i + 1])))
        for i in range(4):                  # decoder: up1..4, skip from cs[3-i]
            cin, cout, skip = cs[4 + i], cs[5 + i], cs[3 - i]
            setattr(self, f"up{i + 1}", nn.ModuleList([
                _Up(be, cin, cout), nn.Sequential(_Res(be, cout + skip, cout), _Res(be, cout, cout))]))
        self.classifier = nn.Sequential(nn.Linear(cs[8], class_num))

    def _init_bn(self):
        for m in self.modules():
            if isinstance(m, nn.BatchNorm1d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)

    def _decode(self, i, y, skip):
        up = getattr(self, f"up{i}")
        return up[1](self.be[0].cat([up[0](y), skip]))


class MinkUNet(_UNetTrunk):
    """network/minkunet.py:14-122: returns (logits [N, class_num], features [N, 96]) row-aligned with the input."""

    def __init__(self, class_num, backend):
        super().__init__(backend, class_num)
        self._init_bn()

    def forward(self, x):
        x0 = self.stem(x)
        x1 = self.stage1(x0)
        x2 = self.stage2(x1)
        x3 = self.stage3(x2)
        x4 = self.stage4(x3)
        y = self._decode(1, x4, x3)
        y = self._decode(2, y, x2)
        y = self._decode(3, y, x1)
        y = self._decode(4, y, x0)
        return self.classifier(y.F), y.F


class SPVCNN(_UNetTrunk):
    """network/spvcnn.py:9-155: the same trunk plus the point branch (three Linear+BN1d+ReLU)."""

    def __init__(self, class_num, backend):
        super().__init__(backend, class_num)
        cs = UNET_CHANNELS
        self.pres = self.vres = 0.05
        self._pv = [PointVoxel(backend)]
        self.point_transforms = nn.ModuleList([
            nn.Sequential(nn.Linear(a, b), nn.BatchNorm1d(b), nn.ReLU(True))
            for a, b in ((cs[0], cs[4]), (cs[4], cs[6]), (cs[6], cs[8]))])
        self._init_bn()
        self.dropout = nn.Dropout(0.3, True)

    def forward(self, x):
        be = self.be[0]
        pv = self._pv[0]
        z = be.PointTensor(x.F, x.C.float())
        x0 = pv.initial_voxelize(z, self.pres, self.vres)
        x0 = self.stem(x0)
        z0 = pv.voxel_to_point(x0, z, nearest=False)
        x1 = self.stage1(pv.point_to_voxel(x0, z0))
        x2 = self.stage2(x1)
        x3 = self.stage3(x2)
        x4 = self.stage4(x3)
        z1 = pv.voxel_to_point(x4, z0)
        z1.F = z1.F + self.point_transforms[0](z0.F)
        y1 = pv.point_to_voxel(x4, z1)
        y1.F = self.dropout(y1.F)
        y1 = self._decode(1, y1, x3)
        y2 = self._decode(2, y1, x2)
        z2 = pv.voxel_to_point(y2, z1)
        z2.F = z2.F + self.point_transforms[1](z1.F)
        y3 = pv.point_to_voxel(y2, z2)
        y3.F = self.dropout(y3.F)
        y3 = self._decode(3, y3, x1)
        y4 = self._decode(4, y3, x0)
        z3 = pv.voxel_to_point(y4, z2)
        z3.F = z3.F + self.point_transforms[2](z2.F)
        return self.classifier(z3.F), z3.F


def seeded_state_dict(template: dict, seed: int = 7122) -> dict:
    """Deterministic, name-keyed random weights: both implementations load the SAME
    dict, never relying on RNG-order equality.  BN running stats are randomised so folding is exercised."""
    out = {}
    for name, t in template.items():
        g = torch.Generator().manual_seed(int.from_bytes(hashlib.sha1(f"{seed}:{name}".encode()).digest()[:4], "little"))
        if name.endswith("num_batches_tracked"):
            out[name] = torch.tensor(100, dtype=t.dtype)
        elif name.endswith("running_mean"):
            out[name] = torch.randn(t.shape, generator=g) * 0.1
        elif name.endswith("running_var"):
            out[name] = torch.rand(t.shape, generator=g) + 0.5
        elif name.endswith("kernel"):
            fan = t.shape[-2] * (t.shape[0] if t.dim() == 3 else 1)
            out[name] = (torch.rand(t.shape, generator=g) * 2 - 1) * (3.0 / fan) ** 0.5
        elif t.dim() == 2:                                   # nn.Linear weight [out, in]
            out[name] = (torch.rand(t.shape, generator=g) * 2 - 1) * (3.0 / t.shape[1]) ** 0.5
        elif name.endswith("weight"):                        # BN gamma
            out[name] = torch.rand(t.shape, generator=g) * 0.5 + 0.75
        else:                                                # biases / BN beta
            out[name] = torch.randn(t.shape, generator=g) * 0.1
        out[name] = out[name].to(t.dtype)
    return out
