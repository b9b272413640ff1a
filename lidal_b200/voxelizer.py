"""Score-mode dataset transform on device: the step right before the hot path.

``dataset/sk_dataset.py:143-169`` + ``collate_fn :188-242`` for the ``inf_reps`` TTA views of one scan: random affine
(+ x flip, yaw), x20, random shift into [0, 8192)^3, ``astype(int)``, ``np.unique(axis=0, return_index, return_inverse)``
with first-point features.  The nine random draws per view stay on the host (same ``RandomState`` call order as the
reference); everything per point runs in lb_* kernels, including the radix-sort based unique.
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib as L

COORD_BITS = 13          # full_scale = 8192 (dataset/sk_dataset.py:56)


def _dbl(values):
    return (C.c_double * len(values))(*[float(v) for v in values])


def voxelize_view(raw_dev: torch.Tensor, rs: np.random.RandomState, batch: int, scale: float = 20.0, full_scale: float = 8192.0,
                  err_out: list | None = None):
    """One augmented view.  raw_dev float32 [Np,4] on device.  Returns coords int32 [Nv,4], feats f32 [Nv,4], inverse int32 [Np]."""
    L.require_cuda(raw_dev)
    raw_dev = raw_dev.contiguous()
    n, dev = raw_dev.shape[0], raw_dev.device
    trans_m = np.eye(3) + rs.randn(3, 3) * 0.1                                      # :144
    trans_m[0][0] *= rs.randint(0, 2) * 2 - 1                                       # :145
    theta = rs.rand() * 2 * math.pi                                                 # :146
    trans_m = np.matmul(trans_m, [[math.cos(theta), math.sin(theta), 0], [-math.sin(theta), math.cos(theta), 0], [0, 0, 1]])
    cp = torch.empty((n, 3), dtype=torch.float64, device=dev)
    feats_p = torch.empty((n, 4), dtype=torch.float32, device=dev)
    L.check(L.lib().lb_tta_transform(L.ptr(raw_dev), n, _dbl(trans_m.reshape(-1)), float(scale), L.ptr(cp), L.ptr(feats_p), L.stream()))
    lo, hi = torch.aminmax(cp, dim=0)                                               # :154-155 (6 doubles to the host, one copy)
    mm = torch.stack((lo, hi)).cpu().numpy()
    cmin, cmax = mm[0], mm[1]
    fs = np.array([full_scale] * 3)
    offset = (-cmin + np.clip(fs - cmax + cmin - 0.001, 0, None) * rs.rand(3)
              + np.clip(fs - cmax + cmin + 0.001, None, 0) * rs.rand(3))            # :156
    coords_p = torch.empty((n, 4), dtype=torch.int, device=dev)
    keys = torch.empty(n, dtype=torch.int64, device=dev)
    err = torch.zeros(1, dtype=torch.int, device=dev)
    L.check(L.lib().lb_tta_quantize(L.ptr(cp), n, _dbl(offset), batch, COORD_BITS, L.ptr(coords_p), L.ptr(keys), L.ptr(err), L.stream()))
    uniq = torch.empty(n, dtype=torch.int64, device=dev)
    from .engine import _counters
    cnt = _counters(dev)                                                            # count lands in pinned host memory
    inverse = torch.empty(n, dtype=torch.int, device=dev)
    first = torch.empty(n, dtype=torch.int, device=dev)
    nbytes = L.lib().lb_unique_ws_bytes(n)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    L.check(L.lib().lb_unique_i64(L.ptr(keys), n, 3 * COORD_BITS, L.ptr(uniq), cnt.ptr(6), L.ptr(inverse), L.ptr(first), L.ptr(ws),
                                  nbytes, L.stream()))
    nv = cnt.read(6)
    if err_out is None:
        assert int(err.item()) == 0, "input voxels are not valid"                   # :160-161
    else:
        err_out.append(err)                                                         # checked once per batch by the caller
    coords_v = torch.empty((nv, 4), dtype=torch.int, device=dev)
    feats_v = torch.empty((nv, 4), dtype=torch.float32, device=dev)
    L.check(L.lib().lb_gather_rows16(L.ptr(coords_p), L.ptr(first), nv, L.ptr(coords_v), L.stream()))
    L.check(L.lib().lb_gather_rows16(L.ptr(feats_p), L.ptr(first), nv, L.ptr(feats_v), L.stream()))
    return coords_v, feats_v, inverse


def tta_batch_gpu(raw_dev: torch.Tensor, seed: int, inf_reps: int = 8):
    """The batch score/prob_inference.py:91-97 consumes, built on device: (coords [N,4], feats [N,4], inverse int64 [reps*Np])."""
    rs = np.random.RandomState(seed)
    coords, feats, inverse, off, errs = [], [], [], 0, []
    for b in range(inf_reps):
        c, f, inv = voxelize_view(raw_dev, rs, b, err_out=errs)
        coords.append(c); feats.append(f); inverse.append(inv.long() + off)
        off += c.shape[0]
    assert int(torch.cat(errs).max().item()) == 0, "input voxels are not valid"             # :160-161
    return torch.cat(coords), torch.cat(feats), torch.cat(inverse)
