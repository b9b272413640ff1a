"""Point<->voxel glue of SPVCNN, bound to a sparse backend.

Behavioural mirror of the reference's ``network/utils.py:13-102`` (the caller side of the
torchsparse boundary): the same sequence of ``F.sphash / sphashquery / spcount / spvoxelize /
calc_ti_weights / spdevoxelize`` calls and the same caching quirks -- ``initial_voxelize`` stores
its query under key ``1`` while later lookups use the stride tuple (``network/utils.py:29-30`` vs
``:39-41``), so the first ``point_to_voxel`` recomputes its query.
"""
from __future__ import annotations

import torch


class PointVoxel:
    def __init__(self, backend):
        self.be = backend
        self.F = backend.nn.functional

    # network/utils.py:13-33
    def initial_voxelize(self, z, init_res, after_res):
        F, be = self.F, self.be
        batch = z.C[:, -1].view(-1, 1)
        scaled = torch.cat([(z.C[:, :3] * init_res) / after_res, batch], 1)
        cell = torch.floor(scaled)
        point_hash = F.sphash(cell.int())
        voxel_hash = torch.unique(point_hash)                   # ascending hash order defines voxel order
        idx_query = F.sphashquery(point_hash, voxel_hash)
        counts = F.spcount(idx_query.int(), len(voxel_hash))
        vox_coords = torch.round(F.spvoxelize(cell, idx_query, counts)).int()
        vox_feats = F.spvoxelize(z.F, idx_query, counts)
        x = be.SparseTensor(vox_feats, vox_coords, 1)
        x.cmaps.setdefault(x.stride, x.coords)
        z.additional_features["idx_query"][1] = idx_query
        z.additional_features["counts"][1] = counts
        z.C = scaled
        return x

    def _cell_hash(self, z, stride, offsets=None):
        s = stride[0]
        cell = torch.cat([torch.floor(z.C[:, :3] / s).int() * s, z.C[:, -1].int().view(-1, 1)], 1)
        return self.F.sphash(cell) if offsets is None else self.F.sphash(cell, offsets)

    # network/utils.py:38-61
    def point_to_voxel(self, x, z):
        F, be = self.F, self.be
        cache = z.additional_features
        if cache is None or cache.get("idx_query") is None or cache["idx_query"].get(x.s) is None:
            idx_query = F.sphashquery(self._cell_hash(z, x.s), F.sphash(x.C))
            counts = F.spcount(idx_query.int(), x.C.shape[0])
            cache["idx_query"][x.s] = idx_query
            cache["counts"][x.s] = counts
        else:
            idx_query, counts = cache["idx_query"][x.s], cache["counts"][x.s]
        out = be.SparseTensor(F.spvoxelize(z.F, idx_query, counts), x.C, x.s)
        out.cmaps, out.kmaps = x.cmaps, x.kmaps
        return out

    # network/utils.py:66-102
    def voxel_to_point(self, x, z, nearest=False):
        F, be = self.F, self.be
        fresh = (z.idx_query is None or z.weights is None or z.idx_query.get(x.s) is None
                 or z.weights.get(x.s) is None)
        if fresh:
            corners = be.nn.utils.get_kernel_offsets(2, x.s, 1, device=z.F.device)
            idx_query = F.sphashquery(self._cell_hash(z, x.s, corners), F.sphash(x.C.to(z.F.device)))
            weights = F.calc_ti_weights(z.C, idx_query, scale=x.s[0]).transpose(0, 1).contiguous()
            idx_query = idx_query.transpose(0, 1).contiguous()
            if nearest:
                weights[:, 1:] = 0.0
                idx_query[:, 1:] = -1
        else:
            idx_query, weights = z.idx_query.get(x.s), z.weights.get(x.s)
        out = be.PointTensor(F.spdevoxelize(x.F, idx_query, weights), z.C, idx_query=z.idx_query, weights=z.weights)
        out.additional_features = z.additional_features
        if fresh:
            out.idx_query[x.s] = idx_query
            out.weights[x.s] = weights
            z.idx_query[x.s] = idx_query
            z.weights[x.s] = weights
        return out
