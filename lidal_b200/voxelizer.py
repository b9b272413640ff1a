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


def draw_view_randoms(rs: np.random.RandomState, inf_reps: int):
    """The random numbers of ``inf_reps`` consecutive score-mode transforms, drawn in the reference's order
    (dataset/sk_dataset.py:144-147,156: randn(3,3), randint, rand, then rand(3), rand(3) per view; none of the draws depends
    on the data).  Returns trans_m [reps,3,3], r1 [reps,3], r2 [reps,3] (float64)."""
    tm, r1, r2 = [], [], []
    for _ in range(inf_reps):
        trans_m = np.eye(3) + rs.randn(3, 3) * 0.1                                  # :144
        trans_m[0][0] *= rs.randint(0, 2) * 2 - 1                                   # :145
        theta = rs.rand() * 2 * math.pi                                             # :146
        trans_m = np.matmul(trans_m, [[math.cos(theta), math.sin(theta), 0], [-math.sin(theta), math.cos(theta), 0], [0, 0, 1]])
        tm.append(trans_m)
        r1.append(rs.rand(3))                                                       # :156, first rand(3)
        r2.append(rs.rand(3))                                                       # :156, second rand(3)
    return np.stack(tm), np.stack(r1), np.stack(r2)


def tta_batch_gpu(raw_dev: torch.Tensor, seed: int, inf_reps: int = 8, scale: float = 20.0, full_scale: float = 8192.0):
    """The batch score/prob_inference.py:91-97 consumes, built on device: (coords int32 [N,4], feats f32 [N,4], inverse int64
    [reps*Np]).  All views go through two launches (lb_tta_views: transform + min/max, shift + quantise) and ONE radix
    unique over (view, x, y, z) keys, which yields the collated order, the first-point rows and the offset inverse indices
    at once; the only host round trip is the voxel count."""
    L.require_cuda(raw_dev)
    raw_dev = raw_dev.contiguous().float()
    n, dev = raw_dev.shape[0], raw_dev.device
    tm, r1, r2 = draw_view_randoms(np.random.RandomState(seed), inf_reps)
    tot = inf_reps * n
    feats_p = torch.empty((tot, 4), dtype=torch.float32, device=dev)
    coords_p = torch.empty((tot, 4), dtype=torch.int, device=dev)
    keys = torch.empty(tot, dtype=torch.int64, device=dev)
    err = torch.zeros(1, dtype=torch.int, device=dev)
    nb = L.lib().lb_tta_views_ws_bytes(n, inf_reps)
    ws = torch.empty(nb, dtype=torch.uint8, device=dev)
    L.check(L.lib().lb_tta_views(L.ptr(raw_dev), n, inf_reps, _dbl(tm.reshape(-1)), _dbl(r1.reshape(-1)), _dbl(r2.reshape(-1)),
                                 float(scale), float(full_scale), COORD_BITS, L.ptr(feats_p), L.ptr(coords_p), L.ptr(keys),
                                 L.ptr(err), L.ptr(ws), nb, L.stream()))
    from .engine import _counters
    cnt = _counters(dev)
    uniq = torch.empty(tot, dtype=torch.int64, device=dev)
    inverse = torch.empty(tot, dtype=torch.int, device=dev)
    first = torch.empty(tot, dtype=torch.int, device=dev)
    nb2 = L.lib().lb_unique_ws_bytes(tot)
    ws2 = torch.empty(nb2, dtype=torch.uint8, device=dev)
    L.check(L.lib().lb_unique_i64(L.ptr(keys), tot, 3 * COORD_BITS + 4, L.ptr(uniq), cnt.ptr(6), L.ptr(inverse), L.ptr(first),
                                  L.ptr(ws2), nb2, L.stream()))
    err_host = cnt.pinned_copy(err, 5)                      # rides on the same event as the count
    nv = cnt.read(6)
    assert int(err_host()) == 0, "input voxels are not valid"                                # :160-161
    coords_v = torch.empty((nv, 4), dtype=torch.int, device=dev)
    feats_v = torch.empty((nv, 4), dtype=torch.float32, device=dev)
    L.check(L.lib().lb_gather_rows16(L.ptr(coords_p), L.ptr(first), nv, L.ptr(coords_v), L.stream()))
    L.check(L.lib().lb_gather_rows16(L.ptr(feats_p), L.ptr(first), nv, L.ptr(feats_v), L.stream()))
    return coords_v, feats_v, inverse.long()


def tta_batch_gpu_per_view(raw_dev: torch.Tensor, seed: int, inf_reps: int = 8):
    """Same result as ``tta_batch_gpu`` through the per-view entry points (one transform / quantise / unique per view)."""
    rs = np.random.RandomState(seed)
    coords, feats, inverse, off, errs = [], [], [], 0, []
    for b in range(inf_reps):
        c, f, inv = voxelize_view(raw_dev, rs, b, err_out=errs)
        coords.append(c); feats.append(f); inverse.append(inv.long() + off)
        off += c.shape[0]
    assert int(torch.cat(errs).max().item()) == 0, "input voxels are not valid"             # :160-161
    return torch.cat(coords), torch.cat(feats), torch.cat(inverse)
