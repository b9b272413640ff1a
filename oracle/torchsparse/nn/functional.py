"""ORACLE: torchsparse/nn/functional/*.py + backend *_cpu.cpp (v1.4.0), restated on torch-CPU/numpy.

Everything is fp32 / int exact arithmetic in the reference's operation order where the
order is observable (conv: offsets ascending, one mm per offset).
"""
import numpy as np
import torch

from ..tensor import SparseTensor
from .utils import get_kernel_offsets, make_ntuple

# Test switch: round the convolution operands (activations, weights, output gradients) to a 16-bit type at exactly the points
# where the CUDA path does (fp32 accumulation stays), so that training-step parity checks the kernels, not the precision choice.
EMULATE_16BIT = None


def _r16(t):
    return t if EMULATE_16BIT is None else t.to(EMULATE_16BIT).float()


FNV_OFFSET = np.uint64(14695981039346656037)
FNV_PRIME = np.uint64(1099511628211)
LOW60 = np.uint64(0x0FFFFFFFFFFFFFFF)


def _fnv(c4: np.ndarray) -> np.ndarray:
    """backend/hash: 64-bit FNV-1a over four int32 taken as uint32, folded to 60 bits."""
    u = c4.astype(np.int32).view(np.uint32).astype(np.uint64)
    h = np.full(u.shape[:-1], FNV_OFFSET, np.uint64)
    with np.errstate(over="ignore"):
        for j in range(4):
            h = (h ^ u[..., j]) * FNV_PRIME
    h = (h >> np.uint64(60)) ^ (h & LOW60)
    return h.astype(np.int64)


def sphash(coords, offsets=None):
    assert coords.dtype == torch.int and coords.dim() == 2 and coords.shape[1] == 4, coords.shape
    c = coords.detach().cpu().numpy()
    if offsets is None:
        return torch.from_numpy(_fnv(c)).to(coords.device)
    assert offsets.dtype == torch.int and offsets.shape[1] == 3
    o = offsets.detach().cpu().numpy()
    q = np.repeat(c[None], o.shape[0], 0).copy()               # [K,N,4]
    q[:, :, :3] += o[:, None, :]
    return torch.from_numpy(_fnv(q)).to(coords.device)          # [K,N]


def sphashquery(queries, references):
    """Index of each query hash inside references, -1 on miss (first insert wins)."""
    q = queries.detach().cpu().numpy().reshape(-1)
    r = references.detach().cpu().numpy().reshape(-1)
    order = np.argsort(r, kind="stable")
    rs = r[order]
    pos = np.searchsorted(rs, q, side="left")
    posc = np.minimum(pos, max(len(rs) - 1, 0))
    hit = (pos < len(rs)) & (rs[posc] == q) if len(rs) else np.zeros_like(q, bool)
    out = np.where(hit, order[posc] if len(rs) else 0, -1).astype(np.int64)
    return torch.from_numpy(out).view(queries.shape).to(queries.device)


def spcount(coords, num):
    idx = coords.detach().cpu().numpy().astype(np.int64)
    idx = idx[idx >= 0]
    return torch.from_numpy(np.bincount(idx, minlength=num).astype(np.int32)).to(coords.device)


class _Voxelize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, idx, counts):
        idx = idx.long()
        ok = idx >= 0
        out = torch.zeros(counts.shape[0], feats.shape[1], dtype=feats.dtype)
        contrib = feats[ok] / counts[idx[ok]].to(feats.dtype).unsqueeze(1)
        out.index_add_(0, idx[ok], contrib)
        ctx.save_for_backward(idx, counts)
        ctx.n = feats.shape[0]
        return out

    @staticmethod
    def backward(ctx, g):
        idx, counts = ctx.saved_tensors
        ok = idx >= 0
        gf = torch.zeros(ctx.n, g.shape[1], dtype=g.dtype)
        gf[ok] = g[idx[ok]] / counts[idx[ok]].to(g.dtype).unsqueeze(1)
        return gf, None, None


def spvoxelize(feats, idx, counts):
    return _Voxelize.apply(feats.contiguous(), idx, counts)


class _Devoxelize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, idx, w):
        idx = idx.long()
        out = torch.zeros(idx.shape[0], feats.shape[1], dtype=feats.dtype)
        for k in range(idx.shape[1]):                       # corner order 0..7, as the backend loop
            ok = idx[:, k] >= 0
            out[ok] += w[ok, k].unsqueeze(1) * feats[idx[ok, k]]
        ctx.save_for_backward(idx, w)
        ctx.m = feats.shape[0]
        return out

    @staticmethod
    def backward(ctx, g):
        idx, w = ctx.saved_tensors
        gf = torch.zeros(ctx.m, g.shape[1], dtype=g.dtype)
        for k in range(idx.shape[1]):
            ok = idx[:, k] >= 0
            gf.index_add_(0, idx[ok, k], w[ok, k].unsqueeze(1) * g[ok])
        return gf, None, None


def spdevoxelize(feats, idx, weights):
    return _Devoxelize.apply(feats.contiguous(), idx, weights)


def calc_ti_weights(coords, idx_query, scale=1):
    with torch.no_grad():
        p = coords
        pf = torch.floor(coords / scale) * scale if scale != 1 else torch.floor(coords)
        pc = pf + scale
        x, y, z = (p[:, i].view(-1, 1) for i in range(3))
        xf, yf, zf = (pf[:, i].view(-1, 1).float() for i in range(3))
        xc, yc, zc = (pc[:, i].view(-1, 1).float() for i in range(3))
        w = torch.cat([
            (xc - x) * (yc - y) * (zc - z), (xc - x) * (yc - y) * (z - zf),
            (xc - x) * (y - yf) * (zc - z), (xc - x) * (y - yf) * (z - zf),
            (x - xf) * (yc - y) * (zc - z), (x - xf) * (yc - y) * (z - zf),
            (x - xf) * (y - yf) * (zc - z), (x - xf) * (y - yf) * (z - zf)], dim=1)
        w = w.transpose(1, 0).contiguous()
        if scale != 1:
            w /= scale ** 3
        w[idx_query == -1] = 0
        w /= torch.sum(w, dim=0) + 1e-8
    return w


def spdownsample(coords, stride=2, kernel_size=2, tensor_stride=1):
    stride, kernel_size, tensor_stride = make_ntuple(stride), make_ntuple(kernel_size), make_ntuple(tensor_stride)
    assert all(stride[k] in (1, kernel_size[k]) for k in range(3)), "only the stride in {1, ks} path is restated"
    ss = torch.tensor([stride[k] * tensor_stride[k] for k in range(3)], dtype=torch.int).unsqueeze(0)
    coords = coords.clone()
    coords[:, :3] = torch.div(coords[:, :3], ss, rounding_mode="floor") * ss
    coords = torch.unique(coords[:, [3, 0, 1, 2]], dim=0)       # sorted by (b, x, y, z)
    return coords[:, [1, 2, 3, 0]].contiguous()


def build_kernel_map(coords, in_stride, kernel_size, stride, dilation):
    """Returns (nbmaps int64 [M,2]=(in,out), nbsizes int64 [K], (n_in, n_out), out_coords, results [K,N_out])."""
    offsets = get_kernel_offsets(kernel_size, stride=in_stride, dilation=dilation)
    references = sphash(coords)
    out_coords = spdownsample(coords, stride, kernel_size, in_stride) if any(s > 1 for s in stride) else coords
    results = sphashquery(sphash(out_coords, offsets), references)          # [K, N_out]
    nbsizes = torch.sum(results != -1, dim=1)
    nbmaps = torch.nonzero(results != -1)
    nbmaps[:, 0] = results.view(-1)[nbmaps[:, 0] * results.size(1) + nbmaps[:, 1]]
    return nbmaps, nbsizes, (coords.shape[0], out_coords.shape[0]), out_coords, results


class _Dense16(torch.autograd.Function):
    """1x1 convolution with the same operand rounding as the CUDA path (only used when EMULATE_16BIT is set)."""

    @staticmethod
    def forward(ctx, feats, weight):
        ctx.save_for_backward(feats, weight)
        return _r16(feats) @ _r16(weight)

    @staticmethod
    def backward(ctx, g):
        feats, weight = ctx.saved_tensors
        g = _r16(g)
        return g @ _r16(weight).t(), _r16(feats).t() @ g


class _Conv(torch.autograd.Function):
    """backend/convolution: per offset gather -> mm -> scatter-add, offsets ascending, fp32."""

    @staticmethod
    def forward(ctx, feats, weight, nbmaps, nbsizes, sizes, transposed):
        out = torch.zeros(sizes[1], weight.shape[-1], dtype=feats.dtype)
        a, b = (1, 0) if transposed else (0, 1)
        cur = 0
        f16, w16 = _r16(feats), _r16(weight)
        for k, n in enumerate(nbsizes.tolist()):
            if n:
                m = nbmaps[cur:cur + n]
                out.index_add_(0, m[:, b], f16[m[:, a]] @ w16[k])
            cur += n
        ctx.save_for_backward(feats, weight, nbmaps, nbsizes)
        ctx.transposed = transposed
        return out

    @staticmethod
    def backward(ctx, g):
        feats, weight, nbmaps, nbsizes = ctx.saved_tensors
        a, b = (1, 0) if ctx.transposed else (0, 1)
        gi, gw = torch.zeros_like(feats), torch.zeros_like(weight)
        cur = 0
        g, f16, w16 = _r16(g), _r16(feats), _r16(weight)
        for k, n in enumerate(nbsizes.tolist()):
            if n:
                m = nbmaps[cur:cur + n]
                go = g[m[:, b]]
                gi.index_add_(0, m[:, a], go @ w16[k].t())
                gw[k] = f16[m[:, a]].t() @ go
            cur += n
        return gi, gw, None, None, None, None


def conv3d(input, weight, kernel_size, bias=None, stride=1, dilation=1, transposed=False):
    feats, coords = input.feats, input.coords
    kernel_size, stride, dilation = make_ntuple(kernel_size), make_ntuple(stride), make_ntuple(dilation)
    if kernel_size == (1, 1, 1) and stride == (1, 1, 1) and dilation == (1, 1, 1):
        feats = feats.matmul(weight) if EMULATE_16BIT is None else _Dense16.apply(feats, weight)
        if bias is not None:
            feats = feats + bias
        output = SparseTensor(feats, coords, input.stride)
    elif not transposed:
        key = (input.stride, kernel_size, stride, dilation)
        kmap = input.kmaps.get(key)
        if kmap is None:
            nbmaps, nbsizes, sizes, out_coords, _ = build_kernel_map(coords, input.stride, kernel_size, stride, dilation)
            kmap = [nbmaps, nbsizes, sizes, out_coords]
            input.kmaps[key] = kmap
        feats = _Conv.apply(feats, weight, kmap[0], kmap[1], kmap[2], False)
        if bias is not None:
            feats = feats + bias
        output = SparseTensor(feats, kmap[3], tuple(input.stride[k] * stride[k] for k in range(3)))
    else:
        ts = tuple(input.stride[k] // stride[k] for k in range(3))
        kmap = input.kmaps[(ts, kernel_size, stride, dilation)]
        feats = _Conv.apply(feats, weight, kmap[0], kmap[1], (kmap[2][1], kmap[2][0]), True)
        if bias is not None:
            feats = feats + bias
        output = SparseTensor(feats, input.cmaps[ts], ts)
    output.cmaps = input.cmaps
    output.cmaps.setdefault(output.stride, output.coords)
    output.kmaps = input.kmaps
    return output
