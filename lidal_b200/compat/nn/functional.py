"""``torchsparse.nn.functional`` surface on the lidal_b200 C ABI.

Each function keeps the name, argument meaning, output shape/dtype and error behaviour of
torchsparse 1.4.0 as used by the reference (network/utils.py:17-25,42-56,69-95) and launches
hand-written sm_100a kernels on the current CUDA stream.  CUDA tensors only.
"""
from __future__ import annotations

import ctypes as C
import weakref

import torch

from ... import _lib as L
from ..tensor import SparseTensor
from .utils import get_kernel_offsets, make_ntuple

__all__ = ["sphash", "sphashquery", "spcount", "spvoxelize", "spdevoxelize", "calc_ti_weights", "spdownsample",
           "conv3d", "build_kernel_map"]

ACT_DTYPE = torch.bfloat16          # 16-bit operand type of the tensor-core convolution in inference (fp32 accumulate)
TRAIN_DTYPE = torch.float16         # operand type when gradients are recorded: fp16's 11-bit significand keeps whole-network
                                    # gradients within cosine 0.99 of the fp32 reference (bf16: ~0.92); output gradients are
                                    # rescaled per tensor by a power of two before the cast, so they stay in fp16's normal range
GRAD_TARGET = 8192.0                # max |g * s| aimed for (fp16 max 65504: headroom for the accumulation of many terms is
                                    # irrelevant -- accumulation is fp32 -- but rounding to fp16 must not overflow)
CONV_TRACE = None                   # set by lidal_b200.profiling to time every conv launch with CUDA events


def set_conv_dtype(dtype):
    global ACT_DTYPE
    assert dtype in (torch.bfloat16, torch.float16)
    ACT_DTYPE = dtype


def _ws(nbytes, device):
    return torch.empty(int(nbytes), dtype=torch.uint8, device=device)


# ------------------------------------------------------------------------------------------- hashing
def sphash(coords, offsets=None):
    assert coords.dtype == torch.int, coords.dtype
    assert coords.dim() == 2 and coords.shape[1] == 4, coords.shape
    L.require_cuda(coords, offsets)
    coords = coords.contiguous()
    n = coords.shape[0]
    if offsets is None:
        out = torch.empty(n, dtype=torch.int64, device=coords.device)
        L.check(L.lib().lb_hash(L.ptr(coords), n, L.ptr(out), L.stream()))
        return out
    assert offsets.dtype == torch.int and offsets.dim() == 2 and offsets.shape[1] == 3, offsets.shape
    offsets = offsets.contiguous()
    k = offsets.shape[0]
    out = torch.empty((k, n), dtype=torch.int64, device=coords.device)
    L.check(L.lib().lb_kernel_hash(L.ptr(coords), n, L.ptr(offsets), k, L.ptr(out), L.stream()))
    return out


def _build_table(keys):
    nbytes = L.lib().lb_hashtable_bytes(keys.numel())
    table = _ws(nbytes, keys.device)
    L.check(L.lib().lb_hashtable_build(L.ptr(keys), keys.numel(), L.ptr(table), nbytes, L.stream()))
    return table, nbytes


def sphashquery(queries, references):
    L.require_cuda(queries, references)
    q = queries.contiguous().view(-1)
    r = references.contiguous().view(-1)
    assert q.dtype == torch.int64 and r.dtype == torch.int64
    table, nbytes = _build_table(r)
    out = torch.empty_like(q)
    L.check(L.lib().lb_hashtable_query(L.ptr(table), nbytes, L.ptr(q), q.numel(), L.ptr(out), L.stream()))
    return out.view(queries.shape)


def spcount(coords, num):
    L.require_cuda(coords)
    idx = coords.contiguous()
    assert idx.dtype == torch.int
    out = torch.empty(int(num), dtype=torch.int, device=idx.device)
    L.check(L.lib().lb_count(L.ptr(idx), idx.numel(), L.ptr(out), int(num), L.stream()))
    return out


# ------------------------------------------------------------------------------------------- point <-> voxel
class _Voxelize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, idx, counts):
        feats = feats.contiguous().float()
        idx = idx.contiguous().int()
        counts = counts.contiguous().int()
        n, c = feats.shape
        m = counts.shape[0]
        out = torch.empty((m, c), dtype=torch.float32, device=feats.device)
        L.check(L.lib().lb_voxelize_fwd(L.ptr(feats), L.ptr(idx), L.ptr(counts), n, m, c, L.ptr(out), L.stream()))
        ctx.save_for_backward(idx, counts)
        ctx.n = n
        return out

    @staticmethod
    def backward(ctx, g):
        idx, counts = ctx.saved_tensors
        g = g.contiguous().float()
        m, c = g.shape
        gf = torch.empty((ctx.n, c), dtype=torch.float32, device=g.device)
        L.check(L.lib().lb_voxelize_bwd(L.ptr(g), L.ptr(idx), L.ptr(counts), ctx.n, m, c, L.ptr(gf), L.stream()))
        return gf, None, None


def spvoxelize(feats, coords, counts):
    L.require_cuda(feats, coords, counts)
    return _Voxelize.apply(feats, coords, counts)


class _Devoxelize(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats, idx, weights):
        feats = feats.contiguous().float()
        idx = idx.contiguous().int()
        weights = weights.contiguous().float()
        n, m, c = idx.shape[0], feats.shape[0], feats.shape[1]
        assert idx.shape[1] == 8 and weights.shape == idx.shape
        out = torch.empty((n, c), dtype=torch.float32, device=feats.device)
        L.check(L.lib().lb_devoxelize_fwd(L.ptr(feats), L.ptr(idx), L.ptr(weights), n, m, c, L.ptr(out), L.stream()))
        ctx.save_for_backward(idx, weights)
        ctx.m = m
        return out

    @staticmethod
    def backward(ctx, g):
        idx, weights = ctx.saved_tensors
        g = g.contiguous().float()
        n, c = g.shape
        gf = torch.empty((ctx.m, c), dtype=torch.float32, device=g.device)
        L.check(L.lib().lb_devoxelize_bwd(L.ptr(g), L.ptr(idx), L.ptr(weights), n, ctx.m, c, L.ptr(gf), L.stream()))
        return gf, None, None


def spdevoxelize(feats, coords, weights):
    L.require_cuda(feats, coords, weights)
    return _Devoxelize.apply(feats, coords, weights)


def calc_ti_weights(coords, idx_query, scale=1):
    L.require_cuda(coords, idx_query)
    with torch.no_grad():
        p = coords.contiguous().float()
        iq = idx_query.contiguous()
        assert iq.dtype == torch.int64 and iq.shape[0] == 8 and iq.shape[1] == p.shape[0]
        out = torch.empty((8, p.shape[0]), dtype=torch.float32, device=p.device)
        L.check(L.lib().lb_ti_weights(L.ptr(p), p.shape[1], L.ptr(iq), p.shape[0], float(scale), L.ptr(out), L.stream()))
    return out


# ------------------------------------------------------------------------------------------- kernel maps
def spdownsample(coords, stride=2, kernel_size=2, tensor_stride=1):
    L.require_cuda(coords)
    stride, kernel_size, tensor_stride = make_ntuple(stride), make_ntuple(kernel_size), make_ntuple(tensor_stride)
    if not all(stride[k] in (1, kernel_size[k]) for k in range(3)):
        raise NotImplementedError("spdownsample: only stride in {1, kernel_size} (all LiDAL layers) is built")
    coords = coords.contiguous()
    n = coords.shape[0]
    ss = (C.c_int * 3)(*[stride[k] * tensor_stride[k] for k in range(3)])
    out = torch.empty_like(coords)
    n_out = torch.zeros(1, dtype=torch.int, device=coords.device)
    nbytes = L.lib().lb_downsample_ws_bytes(n)
    ws = _ws(nbytes, coords.device)
    L.check(L.lib().lb_downsample(L.ptr(coords), n, ss, 15, L.ptr(out), L.ptr(n_out), L.ptr(ws), nbytes, L.stream()))
    m = int(n_out.item())                       # the torchsparse API returns an exactly-sized tensor: one sync
    if m < 0:
        raise L.LidalError("spdownsample: coordinates must lie in [0, 65536) and batch in [0, 32768)")
    return out[:m].contiguous()


class KernelMap:
    """What the reference caches in ``kmaps[(stride, kernel_size, stride, dilation)]``.

    ``nbr``  int32 [K, N_out]: the dense `results` matrix (row of the input voxel or -1) -- what the
             output-stationary kernels consume.
    ``nbmaps`` / ``nbsizes`` (lazy): torchsparse's compacted (in_idx, out_idx) pairs, offset-major.
    ``nbr_t`` (lazy): per-offset inverse for transposed convolution.
    Indexable like torchsparse's list: [0]=nbmaps, [1]=nbsizes, [2]=(n_in, n_out).
    """

    def __init__(self, nbr, n_in, out_coords):
        self.nbr, self.n_in, self.n_out, self.out_coords = nbr, n_in, nbr.shape[1], out_coords
        self._compact = None
        self._nbr_t = None

    def _compacted(self):
        if self._compact is None:
            k, n_out, dev = self.nbr.shape[0], self.n_out, self.nbr.device
            nbmaps = torch.empty((max(k * n_out, 1), 2), dtype=torch.int, device=dev)
            nbsizes = torch.empty(k, dtype=torch.int, device=dev)
            total = torch.zeros(1, dtype=torch.int, device=dev)
            nbytes = L.lib().lb_kmap_compact_ws_bytes(n_out, k)
            ws = _ws(nbytes, dev)
            L.check(L.lib().lb_kmap_compact(L.ptr(self.nbr), n_out, k, L.ptr(nbmaps), L.ptr(nbsizes), L.ptr(total),
                                            L.ptr(ws), nbytes, L.stream()))
            self._compact = (nbmaps[: int(total.item())], nbsizes)
        return self._compact

    @property
    def nbmaps(self):
        return self._compacted()[0]

    @property
    def nbsizes(self):
        return self._compacted()[1]

    @property
    def nbr_t(self):
        if self._nbr_t is None:
            k = self.nbr.shape[0]
            t = torch.empty((k, self.n_in), dtype=torch.int, device=self.nbr.device)
            L.check(L.lib().lb_kmap_transpose(L.ptr(self.nbr), self.n_out, self.n_out, k, L.ptr(t), self.n_in, L.stream()))
            self._nbr_t = t
        return self._nbr_t

    def wgrad_pairs(self, transposed):
        """(pairs int32 [M,2] = (x_row, g_row) offset-major, host prefix offsets) for lb_conv_wgrad, cached per map: the
        offsets need ``nbsizes`` on the host once per map, not once per layer that shares the map."""
        cache = self.__dict__.setdefault("_wgrad", {})
        if transposed not in cache:
            pairs = self.nbmaps
            if transposed:
                pairs = pairs[:, [1, 0]]
            begin = [0]
            for n in self.nbsizes.tolist():
                begin.append(begin[-1] + n)
            cache[transposed] = (pairs.contiguous(), begin)
        return cache[transposed]

    def __getitem__(self, i):
        return (self.nbmaps, self.nbsizes, (self.n_in, self.n_out))[i]


def build_kernel_map(coords, in_stride, kernel_size, stride, dilation) -> KernelMap:
    offsets = get_kernel_offsets(kernel_size, stride=in_stride, dilation=dilation, device=coords.device)
    coords = coords.contiguous()
    table, nbytes = _build_table(sphash(coords))
    out_coords = spdownsample(coords, stride, kernel_size, in_stride) if any(s > 1 for s in stride) else coords
    k, n_out = offsets.shape[0], out_coords.shape[0]
    nbr = torch.empty((k, n_out), dtype=torch.int, device=coords.device)
    L.check(L.lib().lb_kmap_query(L.ptr(table), nbytes, L.ptr(out_coords), n_out, None, L.ptr(offsets), k, L.ptr(nbr),
                                  L.stream()))
    return KernelMap(nbr, coords.shape[0], out_coords)


# ------------------------------------------------------------------------------------------- convolution
_PACKED = {}        # id(Parameter) -> (weakref, version, dtype, packed weight): repack only after an update


def packed_weight_cached(kernel, dtype, transposed=False):
    """Packed weight of an nn.Parameter (``transposed``: of W^T per offset, the dgrad operand), re-packed only when the
    parameter was modified.  Keyed by id() and validated through a weakref, the version counter AND the storage pointer
    (``p.data = t`` / ``p.data.copy_`` do not bump ``_version``; a new storage is caught by the pointer, an in-place
    ``.data`` write is not -- call ``invalidate_packed_weights()`` after such surgery)."""
    key = (id(kernel), bool(transposed))
    hit = _PACKED.get(key)
    if hit is not None and hit[0]() is kernel and hit[1] == kernel._version and hit[2] == dtype and hit[3].device == kernel.device \
            and hit[4] == kernel.data_ptr():
        return hit[3]
    if transposed:
        w3 = kernel if kernel.dim() == 3 else kernel.unsqueeze(0)
        packed = pack_weight(w3.transpose(1, 2), dtype)
    else:
        packed = pack_weight(kernel, dtype)
    if len(_PACKED) > 4096:
        for k_ in [k for k, v in _PACKED.items() if v[0]() is None]:
            del _PACKED[k_]
    try:
        _PACKED[key] = (weakref.ref(kernel), kernel._version, dtype, packed, kernel.data_ptr())
    except TypeError:
        pass
    return packed


def invalidate_packed_weights():
    """Forget every cached packed weight (after writing parameters through ``.data``)."""
    _PACKED.clear()


def pack_weight(kernel, dtype):
    """fp32 [K, Cin, Cout] (or [Cin, Cout]) -> 16-bit [K, Cout, Cin] for the implicit GEMM."""
    w = kernel.detach().contiguous().float()
    if w.dim() == 2:
        w = w.unsqueeze(0)
    k, cin, cout = w.shape
    packed = torch.empty((k, cout, cin), dtype=dtype, device=w.device)
    L.check(L.lib().lb_conv_pack_weight(L.ptr(w), k, cin, cout, L.DT_OF[dtype], L.ptr(packed), L.stream()))
    return packed


def conv_forward(feats16, packed_w, nbr, n_out, *, scale=None, shift=None, residual=None, relu=False,
                 out=None, out_dtype=torch.float32, force_simt=False, extra_flags=0):
    """One lb_conv_fwd launch.  feats16 [n_in, c_in] 16-bit (may be a column slice); nbr int32 [K, >=n_out] or None."""
    k, cout, cin = packed_w.shape
    assert feats16.dtype == packed_w.dtype and feats16.stride(1) == 1
    if out is None:
        out = torch.empty((n_out, cout), dtype=out_dtype, device=feats16.device)
    a = L.ConvArgs()
    a.inp, a.n_in, a.ld_in = feats16.data_ptr(), feats16.shape[0], feats16.stride(0)
    a.out, a.n_out, a.ld_out = out.data_ptr(), n_out, out.stride(0)
    a.n_out_dev = None
    a.nbr, a.nbr_ld = (nbr.data_ptr(), nbr.stride(0)) if nbr is not None else (None, 0)
    a.out_rows = None
    a.weight, a.k_vol, a.c_in, a.c_out = packed_w.data_ptr(), k, cin, cout
    a.scale = scale.data_ptr() if scale is not None else None
    a.shift = shift.data_ptr() if shift is not None else None
    a.residual, a.ld_res = (residual.data_ptr(), residual.stride(0)) if residual is not None else (None, 0)
    a.act_dtype, a.out_dtype = L.DT_OF[feats16.dtype], L.DT_OF[out.dtype]
    a.flags = (L.LB_CONV_RELU if relu else 0) | (L.LB_CONV_FORCE_SIMT if force_simt else 0) | extra_flags
    a.sched_ws = L.conv_sched_ws()
    if CONV_TRACE is None:
        L.check(L.lib().lb_conv_fwd(C.byref(a), L.stream()))
        return out
    ev = CONV_TRACE.begin()
    L.check(L.lib().lb_conv_fwd(C.byref(a), L.stream()))
    tc = (not force_simt) and L.lib().lb_conv_uses_tensor_cores(k, cin, cout, a.act_dtype) and feats16.stride(0) % 8 == 0
    CONV_TRACE.end(ev, nbr, n_out, k, cin, cout, tc)
    return out


def _to16(x, dtype):
    if x.dtype == dtype:
        return x.contiguous()
    x = x.contiguous()
    out = torch.empty(x.shape, dtype=dtype, device=x.device)
    L.check(L.lib().lb_cast(L.ptr(x), L.DT_OF[x.dtype], x.stride(0), L.ptr(out), L.DT_OF[dtype], out.stride(0),
                            x.shape[0], x.shape[1], L.stream()))
    return out


def _scaled16(g, dtype):
    """fp32 gradient -> (16-bit g * s, inv_vec float[256] = 1/s, scale float[2] = {s, 1/s}); s is a power of two chosen on
    device from max |g| (no host round trip)."""
    g = g.float()
    if g.stride(1) != 1:
        g = g.contiguous()
    rows, cols = g.shape
    dev = g.device
    amax = torch.empty(1, dtype=torch.float32, device=dev)
    L.check(L.lib().lb_absmax_f32(L.ptr(g), g.stride(0), rows, cols, L.ptr(amax), L.stream()))
    g16 = torch.empty((rows, cols), dtype=dtype, device=dev)
    scale = torch.empty(2, dtype=torch.float32, device=dev)
    inv_vec = torch.empty(256, dtype=torch.float32, device=dev)
    L.check(L.lib().lb_cast_scaled(L.ptr(g), g.stride(0), L.ptr(g16), L.DT_OF[dtype], cols, rows, cols, L.ptr(amax), GRAD_TARGET,
                                   L.ptr(scale), L.ptr(inv_vec), L.stream()))
    return g16, inv_vec, scale


class _ConvFunction(torch.autograd.Function):
    """Forward on the implicit-GEMM kernel; dgrad reuses it with swapped map roles and W^T; wgrad is one
    tcgen05 split-K contraction over the map pairs (lb_conv_wgrad).  When gradients are recorded the operands are fp16
    (TRAIN_DTYPE) and the 16-bit copy of the input is what is saved for wgrad."""

    @staticmethod
    def forward(ctx, feats, kernel, nbr, n_out, kmap, transposed):
        train = any(ctx.needs_input_grad[:2])
        dt = TRAIN_DTYPE if train else ACT_DTYPE
        w = packed_weight_cached(kernel, dt)
        x16 = _to16(feats.float(), dt)
        out = conv_forward(x16, w, nbr, n_out)
        ctx.save_for_backward(x16, kernel)
        ctx.kmap, ctx.transposed = kmap, transposed
        return out

    @staticmethod
    def backward(ctx, g):
        x16, kernel = ctx.saved_tensors
        kmap, transposed = ctx.kmap, ctx.transposed
        dt = x16.dtype
        w3 = kernel if kernel.dim() == 3 else kernel.unsqueeze(0)
        gi = gw = None
        g16, inv_vec, scale = _scaled16(g, dt)
        if ctx.needs_input_grad[0]:
            if kmap is None:
                nbr_back = None
            else:
                nbr_back = kmap.nbr if transposed else kmap.nbr_t
            wt = packed_weight_cached(kernel, dt, transposed=True)          # packed [K, Cin (as output), Cout (as input)]
            cin = wt.shape[1]
            if cin <= 256:
                gi = conv_forward(g16, wt, nbr_back, x16.shape[0], scale=inv_vec)
            else:                                                           # MMA N <= 256: produce the input grad in column slices
                gi = torch.empty((x16.shape[0], cin), dtype=torch.float32, device=g.device)
                for c0 in range(0, cin, 128):
                    c1 = min(c0 + 128, cin)
                    conv_forward(g16, wt[:, c0:c1, :].contiguous(), nbr_back, x16.shape[0], out=gi[:, c0:c1], scale=inv_vec)
        if ctx.needs_input_grad[1]:
            gw = _wgrad(x16, g16, kmap, transposed, w3.shape)
            gw = (gw * scale[1]).view(kernel.shape)
        return gi, gw, None, None, None, None


def _wgrad(x16, g16, kmap, transposed, wshape):
    """grad of the `kernel` parameter [K, Cin, Cout] from 16-bit operands.  Tensor-core kernel (lb_conv_wgrad) whenever
    Cout % 32 == 0; the input is zero-padded to a multiple of 8 channels if needed.  Other shapes use one gathered GEMM per offset."""
    k_vol, cin, cout = wshape
    dev = x16.device
    if cout % 32 == 0 and cout <= 256 and k_vol <= 27:
        if kmap is None:
            idx = torch.arange(x16.shape[0], dtype=torch.int, device=dev)
            pairs = torch.stack([idx, idx], 1).contiguous()
            begin = [0, x16.shape[0]]
        else:
            pairs, begin = kmap.wgrad_pairs(transposed)
        cin_p = (cin + 7) // 8 * 8
        if cin_p != cin:
            xp = torch.zeros((x16.shape[0], cin_p), dtype=x16.dtype, device=dev)
            xp[:, :cin] = x16
            x16 = xp
        gw = torch.empty((k_vol, cin_p, cout), dtype=torch.float32, device=dev)
        pb = (C.c_int * len(begin))(*begin)
        L.check(L.lib().lb_conv_wgrad(L.ptr(x16), x16.shape[0], x16.stride(0), L.ptr(g16), g16.shape[0], g16.stride(0),
                                      L.ptr(pairs), pb, k_vol, cin_p, cout, L.DT_OF[x16.dtype], L.ptr(gw), L.stream()))
        return gw[:, :cin, :] if cin_p != cin else gw
    gw = torch.zeros(wshape, dtype=torch.float32, device=dev)
    if kmap is None:
        gw[0] = x16.float().t() @ g16.float()
        return gw
    nbmaps, nbsizes = kmap.nbmaps.long(), kmap.nbsizes.tolist()
    a, b = (1, 0) if transposed else (0, 1)
    cur = 0
    for k, n in enumerate(nbsizes):
        if n:
            m = nbmaps[cur:cur + n]
            gw[k] = x16[m[:, a]].float().t() @ g16[m[:, b]].float()
        cur += n
    return gw


def conv3d(input, weight, kernel_size, bias=None, stride=1, dilation=1, transposed=False):
    feats, coords = input.feats, input.coords
    L.require_cuda(feats, coords, weight)
    kernel_size, stride, dilation = make_ntuple(kernel_size), make_ntuple(stride), make_ntuple(dilation)
    if kernel_size == (1, 1, 1) and stride == (1, 1, 1) and dilation == (1, 1, 1):
        out_feats = _ConvFunction.apply(feats, weight, None, feats.shape[0], None, False)
        output = SparseTensor(out_feats, coords, input.stride)
    elif not transposed:
        key = (input.stride, kernel_size, stride, dilation)
        kmap = input.kmaps.get(key)
        if kmap is None:
            kmap = build_kernel_map(coords, input.stride, kernel_size, stride, dilation)
            input.kmaps[key] = kmap
        out_feats = _ConvFunction.apply(feats, weight, kmap.nbr, kmap.n_out, kmap, False)
        output = SparseTensor(out_feats, kmap.out_coords, tuple(input.stride[k] * stride[k] for k in range(3)))
    else:
        ts = tuple(input.stride[k] // stride[k] for k in range(3))
        kmap = input.kmaps[(ts, kernel_size, stride, dilation)]
        out_feats = _ConvFunction.apply(feats, weight, kmap.nbr_t, kmap.n_in, kmap, True)
        output = SparseTensor(out_feats, input.cmaps[ts], ts)
    if bias is not None:
        output.feats = output.feats + bias
    output.cmaps = input.cmaps
    output.cmaps.setdefault(output.stride, output.coords)
    output.kmaps = input.kmaps
    return output
