"""Fused inference plan for LiDAL's two networks (eval mode), built from any module with the reference's layout
(network/minkunet.py:14-122, network/spvcnn.py:9-155 -- or the mirrors in lidal_b200.network).

What changes relative to running the modules one by one through ``lidal_b200.compat``:
  * eval-mode BatchNorm, ReLU and the residual add are folded into the convolution epilogue (one kernel per conv,
    one HBM round trip per activation instead of four);
  * activations stay 16-bit (bf16 default) between convolutions, fp32 accumulate in TMEM;
  * ``torchsparse.cat`` is free: producers write straight into column slices of the concatenated buffer;
  * packed weights and folded scale/shift are prepared once at construction;
  * SPVCNN's point branch uses fused query kernels (floor/hash/lookup/trilinear weights in one pass) and 16-bit
    gathers; ``Linear+BN1d+ReLU`` point MLPs run on the same tcgen05 kernel as 1x1 convolutions with the
    devoxelised features as the epilogue's residual.
Arithmetic is otherwise the reference's: same kernel maps, same voxel order, same offset order.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L
from .compat.nn import functional as F
from .compat.nn.utils import get_kernel_offsets


# The TMA gather producer of lb_conv_fwd (LIDAL_TMA_GATHER=1, off by default: measured slower than cp.async in situ) fetches
# the rows of missing neighbours from a pool of all-zero rows behind the input (lb_conv_args.in_pad_rows).
ZPAD = int(__import__('os').environ.get('LIDAL_ZPAD', 64)) if int(__import__('os').environ.get('LIDAL_TMA_GATHER', 0)) else 0


def _alloc(n, c, dtype, device):
    """Activation buffer [n, c]; with the TMA gather producer enabled it is backed by n + ZPAD rows whose tail is zero."""
    if ZPAD == 0:
        return torch.empty((n, c), dtype=dtype, device=device)
    buf = torch.empty((n + ZPAD, c), dtype=dtype, device=device)
    buf[n:].zero_()
    buf._zpad = True
    return buf[:n]


def _pad_rows(x):
    """ZPAD if ``x`` (a row-0 view of an ``_alloc`` buffer, possibly a column slice) is followed by the zero pool, else 0."""
    if ZPAD == 0:
        return 0
    base = x._base if x._base is not None else x
    if getattr(base, "_zpad", False) and base.dim() == 2 and base.shape[0] == x.shape[0] + ZPAD and x.stride(0) == base.stride(0) \
            and x.storage_offset() - base.storage_offset() < base.stride(0):
        return ZPAD
    return 0


def _fold_bn(bn, extra_bias=None):
    scale = (bn.weight / torch.sqrt(bn.running_var + bn.eps)).float()
    shift = (bn.bias - bn.running_mean * scale).float()
    if extra_bias is not None:
        shift = shift + extra_bias.float() * scale
    return scale.contiguous(), shift.contiguous()


class _Conv:
    """One fused launch: packed 16-bit weight [K, Cout, Cin] + per-channel scale/shift + flags."""

    def __init__(self, kernel, bn=None, relu=False, dtype=torch.bfloat16, bias=None, pad_out_to=None, pack8=False):
        w = kernel.detach()
        if w.dim() == 2:
            w = w.unsqueeze(0)
        self.k, self.cin, self.cout = w.shape
        self.pack8 = pack8
        self.cout_real = self.cout
        dev = w.device
        if bn is not None:
            self.scale, self.shift = (t.detach().to(dev) for t in _fold_bn(bn, bias))
        else:
            self.scale = None
            self.shift = bias.detach().float().contiguous() if bias is not None else None
        if pad_out_to and self.cout % pad_out_to:
            pad = pad_out_to - self.cout % pad_out_to
            w = torch.cat([w, torch.zeros(self.k, self.cin, pad, device=dev, dtype=w.dtype)], 2)
            if self.shift is not None:
                self.shift = torch.cat([self.shift, torch.zeros(pad, device=dev)])
            if self.scale is not None:
                self.scale = torch.cat([self.scale, torch.ones(pad, device=dev)])
            self.cout += pad
        if pack8:
            # LB_CONV_PACK8: [Cout][kpad] with column = offset * 8 + channel (channels zero-padded to 8)
            assert self.cin <= 8
            kpad = (self.k * 8 + 63) // 64 * 64
            w8 = torch.zeros(kpad // 8, 8, self.cout, device=dev, dtype=torch.float32)
            w8[: self.k, : self.cin] = w.float()
            self.w = w8.permute(2, 0, 1).reshape(1, self.cout, kpad).to(dtype).contiguous()
            self.cin = 8
        else:
            self.w = F.pack_weight(w.float(), dtype)
        self.relu = relu

    def __call__(self, x, nbr, n_out, out=None, residual=None, out_dtype=None, relu_first=False, no_lean=False):
        tile_masks = getattr(nbr, "tile_masks", None)
        nbr, out_rows = nbr if isinstance(nbr, tuple) else (nbr, None)
        out_dtype = out_dtype or x.dtype
        if out is None:
            out = _alloc(n_out, self.cout, out_dtype, x.device) if out_dtype != torch.float32 else \
                torch.empty((n_out, self.cout), dtype=out_dtype, device=x.device)
        a = L.ConvArgs()
        a.inp, a.n_in, a.ld_in = x.data_ptr(), x.shape[0], x.stride(0)
        a.out, a.n_out, a.ld_out = out.data_ptr(), n_out, out.stride(0)
        a.n_out_dev = None
        a.nbr, a.nbr_ld = (nbr.data_ptr(), nbr.stride(0)) if nbr is not None else (None, 0)
        a.out_rows = out_rows.data_ptr() if out_rows is not None else None
        a.weight, a.k_vol, a.c_in, a.c_out = self.w.data_ptr(), self.k, self.cin, self.cout
        a.scale = self.scale.data_ptr() if self.scale is not None else None
        a.shift = self.shift.data_ptr() if self.shift is not None else None
        a.residual, a.ld_res = (residual.data_ptr(), residual.stride(0)) if residual is not None else (None, 0)
        a.act_dtype, a.out_dtype = L.DT_OF[x.dtype], L.DT_OF[out.dtype]
        a.flags = ((L.LB_CONV_RELU if self.relu else 0) | (L.LB_CONV_RELU_FIRST if relu_first else 0)
                   | (L.LB_CONV_PACK8 if self.pack8 else 0) | (L.LB_CONV_TILE128 if FORCE_TILE128 else 0)
                   | (L.LB_CONV_NO_STAGED if NO_STAGED else 0) | (L.LB_CONV_NO_LEAN if no_lean else 0))
        a.sched_ws = L.conv_sched_ws()
        a.in_pad_rows = _pad_rows(x) if nbr is not None else 0
        a.tile_masks = tile_masks.data_ptr() if tile_masks is not None else None
        trace = F.CONV_TRACE
        ev = trace.begin() if trace is not None else None
        L.check(L.lib().lb_conv_fwd(C.byref(a), L.stream()))
        if trace is not None:
            tc = self.pack8 or (L.lib().lb_conv_uses_tensor_cores(self.k, self.cin, self.cout, a.act_dtype) and x.stride(0) % 8 == 0)
            trace.end(ev, nbr, n_out, self.k, self.cin, self.cout, tc)
        return out


class _Res:
    def __init__(self, block, dtype):
        net = block.net
        self.c1 = _Conv(net[0].kernel, net[1], relu=True, dtype=dtype)
        self.c2 = _Conv(net[3].kernel, net[4], relu=True, dtype=dtype)        # ReLU after the residual add
        ds = block.downsample
        self.skip = None if isinstance(ds, torch.nn.Identity) else _Conv(ds[0].kernel, ds[1], relu=False, dtype=dtype)

    def __call__(self, x, nbr, n, out=None):
        t = self.c1(x, nbr, n)
        s = x if self.skip is None else self.skip(x, None, n)
        return self.c2(t, nbr, n, out=out, residual=s)


_OFFSETS = {}
import os as _os
import time as _time
FORCE_TILE128 = bool(int(_os.environ.get('LIDAL_TILE128', '0')))   # A/B switch: 128-row CTA tiles everywhere
NO_STAGED = bool(int(_os.environ.get('LIDAL_NO_STAGED', '0')))       # A/B switch: per-thread epilogue stores
SORT_MAPS = True        # group rows by neighbour mask (tile-level offset skipping); False = natural row order
RESERVE_FACTOR = float(_os.environ.get('LIDAL_PIPE_RESERVE', '6'))   # side-stream pool parked at first use, in units of one prepare()'s allocations
PREP_PRIORITY = int(_os.environ.get('LIDAL_PREP_PRIORITY', '-1'))   # CUDA stream priority of the map-construction stream (A/B: 0 = default)
SORT_DN = bool(int(_os.environ.get('LIDAL_SORT_DN', '0')))          # A/B switch: mask-sort the strided (k = 8) maps as well
TILE_MASKS = bool(int(_os.environ.get('LIDAL_TILE_MASKS', '1')))    # A/B switch: per-tile offset masks (prologue-free conv producer)


def _offsets(ks, stride, dev):
    key = (ks, stride, str(dev))
    if key not in _OFFSETS:
        _OFFSETS[key] = get_kernel_offsets(ks, stride, 1, device=dev)
    return _OFFSETS[key]


class _HostCounters:
    """Row counts that kernels write straight into pinned host memory (the pointer is device-visible under UVA).  Reading
    one costs an event wait -- not a 4-byte D2H copy, which would queue on the copy engine behind the logits download of
    the previous step (HostPipeline) and stall map construction for milliseconds."""

    def __init__(self, slots: int = 16):
        self.buf = torch.zeros(slots, dtype=torch.int32).pin_memory()
        self.ev = torch.cuda.Event()

    def ptr(self, i):
        return C.c_void_p(self.buf.data_ptr() + 4 * i)

    def read(self, i, ev=None) -> int:
        """Value of slot i.  ``ev`` (from ``mark()``, recorded right after the producing launch) makes the wait cover only
        that launch: work queued afterwards keeps the GPU busy while the host learns the count."""
        if ev is None:
            ev = self.ev
            ev.record()
        ev.synchronize()
        return int(self.buf[i])

    def mark(self):
        ev = torch.cuda.Event()
        ev.record()
        return ev

    def pinned_copy(self, dev_scalar: torch.Tensor, i: int):
        """Queue a copy of a device int32 scalar into slot i; the returned callable reads it after the next ``read``."""
        self.buf[i:i + 1].copy_(dev_scalar.reshape(1), non_blocking=True)
        return lambda: int(self.buf[i])


class StageTrace:
    """Optional per-stage device timing for bench.py's ``roofline_by_stage``: CUDA events on the launching stream around
    each non-conv stage of a step, with the stage's ALGORITHMIC bytes (SURVEY.md section 8d formulas)."""

    def __init__(self):
        self.records = []          # (name, ev0, ev1, bytes)

    def summarise(self):
        torch.cuda.synchronize()
        out: dict = {}
        for name, e0, e1, nbytes in self.records:
            d = out.setdefault(name, {"ms": 0.0, "bytes": 0, "calls": 0})
            d["ms"] += e0.elapsed_time(e1)
            d["bytes"] += nbytes
            d["calls"] += 1
        return out


STAGE_TRACE: StageTrace | None = None


class _stage:
    """``with _stage(name, bytes):`` -- free when tracing is off."""
    __slots__ = ("name", "nbytes", "ev")

    def __init__(self, name, nbytes):
        self.name, self.nbytes, self.ev = name, nbytes, None

    def __enter__(self):
        if STAGE_TRACE is not None:
            self.ev = torch.cuda.Event(enable_timing=True)
            self.ev.record()
        return self

    def __exit__(self, *exc):
        if self.ev is not None and STAGE_TRACE is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            STAGE_TRACE.records.append((self.name, self.ev, e1, int(self.nbytes() if callable(self.nbytes) else self.nbytes)))
        return False


_COUNTERS = {}


def _counters(dev) -> _HostCounters:
    key = str(dev)
    if key not in _COUNTERS:
        _COUNTERS[key] = _HostCounters()
    return _COUNTERS[key]


class SortedMap(tuple):
    """(nbr_sorted, perm) of a mask-sorted kernel map, plus ``tile_masks``: the active-offset mask of every 128-row group
    (lb_kmap_tile_masks), which lets lb_conv_fwd run without its per-tile prologue."""

    def __new__(cls, table, perm, tile_masks=None):
        self = super().__new__(cls, (table, perm))
        self.tile_masks = tile_masks
        return self


def tile_masks_of(table):
    """uint32 (stored as int32) [ceil(n / 128)]: bit k = some row of the 128-row group has a neighbour at offset k."""
    k, n = table.shape
    tm = torch.empty(max((n + 127) // 128, 1), dtype=torch.int, device=table.device)
    L.check(L.lib().lb_kmap_tile_masks(L.ptr(table), table.stride(0), n, k, L.ptr(tm), L.stream()))
    return tm


def _mask_sorted(nbr):
    """(nbr_sorted, perm): rows grouped by neighbour mask so 128-row tiles skip absent offsets (lb_kmap_sort_by_mask)."""
    k, n = nbr.shape
    perm = torch.empty(n, dtype=torch.int, device=nbr.device)
    ld = (n + 3) // 4 * 4                # 16-byte aligned table rows: the conv kernel reads them with 128-bit loads
    out = torch.empty((k, ld), dtype=torch.int, device=nbr.device)[:, :n]
    nbytes = L.lib().lb_kmap_sort_ws_bytes(n)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=nbr.device)
    tm = torch.empty(max((n + 127) // 128, 1), dtype=torch.int, device=nbr.device) if TILE_MASKS else None
    L.check(L.lib().lb_kmap_sort_by_mask_tm(L.ptr(nbr), nbr.stride(0), n, k, L.ptr(perm), L.ptr(out), ld,
                                            L.ptr(tm) if tm is not None else None, L.ptr(ws), nbytes, L.stream()))
    return SortedMap(out, perm, tm)


def queue_level_counts(cell_coords, cnt, slot):
    """Queue lb_level_counts on int32 [n,4] cell coordinates (points or voxels); the 4 coarse-level row counts land in the
    pinned slots [slot, slot + 4) of ``cnt`` and are valid after the next ``cnt.read``."""
    n = cell_coords.shape[0]
    nbytes = L.lib().lb_level_counts_ws_bytes(n, 4)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=cell_coords.device)
    L.check(L.lib().lb_level_counts(L.ptr(cell_coords), n, 4, cnt.ptr(slot), L.ptr(ws), nbytes, L.stream()))
    return ws


class Maps:
    """Coordinates and neighbour tables of the 5 resolution levels (9 kernel maps) for one batch.  Every table is kept
    as (mask-sorted table, row permutation): the conv kernel walks rows in permuted order and scatters through out_rows."""

    def __init__(self, coords, level_counts=None):
        """``level_counts``: row counts of levels 1..4 if the caller already knows them (``level_counts_of``); otherwise
        they are fetched here in ONE host round trip -- the construction below is then free of host synchronisation."""
        self.coords, self.n, self.nbr3, self.nbr_dn, self.nbr_up, self.tables = [coords], [coords.shape[0]], [], [], [], []
        dev = coords.device
        if level_counts is None:
            cnt = _counters(dev)
            queue_level_counts(coords, cnt, 8)
            cnt.read(8)
            level_counts = [int(cnt.buf[8 + l]) for l in range(4)]
        for lvl in range(5):
            s = 2 ** lvl
            c = self.coords[lvl]
            n_c = c.shape[0]
            if lvl < 4:
                # level transition first (no sort, no hash queries): parents in first-occurrence order + both maps.  Its row
                # count is the only thing the host needs; everything queued below keeps the GPU busy while it is read back.
                m_c = level_counts[lvl]              # known up front (lb_level_counts): no host round trip per level
                cn_full = torch.empty_like(c)        # the ABI sizes its outputs by the upper bound n_c (it learns n_out on the device)
                cnt = _counters(dev)
                ld_dn = (n_c + 3) // 4 * 4          # 16-byte aligned table rows (128-bit index loads in the conv kernel)
                dn_full = torch.empty((8, ld_dn), dtype=torch.int, device=dev)
                up = torch.empty((8, n_c), dtype=torch.int, device=dev)
                nbytes = L.lib().lb_downsample_maps_ws_bytes(n_c)
                ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
                L.check(L.lib().lb_downsample_maps(L.ptr(c), n_c, s, L.ptr(cn_full), cnt.ptr(lvl), L.ptr(dn_full), ld_dn, L.ptr(up),
                                                   L.ptr(ws), nbytes, L.stream()))
            table = F._build_table(F.sphash(c))
            self.tables.append(table)
            off3 = _offsets(3, s, dev)
            nbr3 = torch.empty((27, n_c), dtype=torch.int, device=dev)
            # submanifold map: in == out coordinates, point-symmetric offsets -> probe 14 offsets, mirror the other 13
            L.check(L.lib().lb_kmap_query_sym(L.ptr(table[0]), table[1], L.ptr(c), n_c, L.ptr(off3), 27,
                                              L.ptr(nbr3), nbr3.stride(0), L.stream()))
            self.nbr3.append(_mask_sorted(nbr3) if SORT_MAPS else nbr3)
            if lvl == 4:
                break
            self.nbr_up.append(_mask_sorted(up) if SORT_MAPS else up)
            cn = cn_full[:m_c]
            dn = dn_full[:, :m_c]
            self.coords.append(cn)
            self.n.append(cn.shape[0])
            # strided (k = 8) maps: every 128-row tile of parents has nearly all 8 children offsets anyway; sorting them
            # costs 8 launches per level on the latency-bound side stream and saves ~0.03 ms of convolution per step
            # (measured: 1070 -> 1107 scans/s without the sort).  They still get tile masks for the short prologue.
            self.nbr_dn.append(_mask_sorted(dn) if (SORT_MAPS and SORT_DN) else
                               (SortedMap(dn, None, tile_masks_of(dn)) if TILE_MASKS and dn.stride(0) % 4 == 0 else dn))


def _maps_algorithmic_bytes(self):
    total = 0
    for lvl in range(5):
        n = self.n[lvl]
        total += 16 * n + 16 * n + 4 * 27 * n + 24 * n                 # submanifold k3 map + its hash table
        if lvl < 4:
            total += 16 * n + 16 * self.n[lvl + 1] + 2 * 4 * 8 * n        # k2s2 map pair (down + transposed)
    return total


Maps.algorithmic_bytes = _maps_algorithmic_bytes


class InferenceEngine:
    def __init__(self, model, dtype=torch.bfloat16):
        assert not model.training, "InferenceEngine folds BatchNorm: call model.eval() first"
        self.dtype = dtype
        self.is_spvcnn = hasattr(model, "point_transforms")
        st = model.stem
        self.stem0 = _Conv(st[0].kernel, st[1], relu=True, dtype=dtype, pack8=True)
        self.stem1 = _Conv(st[3].kernel, st[4], relu=True, dtype=dtype)
        self.down, self.enc = [], []
        for i in range(1, 5):
            stage = getattr(model, f"stage{i}")
            self.down.append(_Conv(stage[0].net[0].kernel, stage[0].net[1], relu=True, dtype=dtype))
            self.enc.append((_Res(stage[1], dtype), _Res(stage[2], dtype)))
        self.up, self.dec = [], []
        for i in range(1, 5):
            up = getattr(model, f"up{i}")
            self.up.append(_Conv(up[0].net[0].kernel, up[0].net[1], relu=True, dtype=dtype))
            self.dec.append((_Res(up[1][0], dtype), _Res(up[1][1], dtype)))
        lin = model.classifier[0]
        self.n_cls = lin.out_features
        self.classifier = _Conv(lin.weight.detach().t().contiguous(), None, relu=False, dtype=dtype, bias=lin.bias,
                                pad_out_to=32)
        if self.is_spvcnn:
            self.mlp = [_Conv(seq[0].weight.detach().t().contiguous(), seq[1], relu=True, dtype=dtype, bias=seq[0].bias)
                        for seq in model.point_transforms]
            self.pres, self.vres = model.pres, model.vres
        self.device = lin.weight.device

    # ---- trunk pieces
    def _cast(self, x, dtype):
        out = _alloc(x.shape[0], x.shape[1], dtype, x.device)
        L.check(L.lib().lb_cast(L.ptr(x), L.DT_OF[x.dtype], x.stride(0), L.ptr(out), L.DT_OF[dtype], out.stride(0),
                                x.shape[0], x.shape[1], L.stream()))
        return out

    def _pad8(self, feats):
        """fp32 [N, <=8] -> 16-bit [N, 8] zero padded (16-byte rows for the PACK8 stem conv)."""
        feats = feats.contiguous()
        out = torch.zeros((feats.shape[0], 8), dtype=self.dtype, device=feats.device)
        L.check(L.lib().lb_cast(L.ptr(feats), L.DT_OF[feats.dtype], feats.stride(0), L.ptr(out), L.DT_OF[self.dtype], 8,
                                feats.shape[0], feats.shape[1], L.stream()))
        return out

    def _stem(self, feats16, m, cat4):
        a = self.stem0(feats16, m.nbr3[0], m.n[0])
        return self.stem1(a, m.nbr3[0], m.n[0], out=cat4[:, self.up[3].cout:])

    def _encode(self, x, m, lvl, out=None):
        """stage `lvl` (1..4): strided conv from level lvl-1, two residual blocks."""
        d = self.down[lvl - 1](x, m.nbr_dn[lvl - 1], m.n[lvl])
        r = self.enc[lvl - 1][0](d, m.nbr3[lvl], m.n[lvl])
        return self.enc[lvl - 1][1](r, m.nbr3[lvl], m.n[lvl], out=out)

    def _decode(self, y, m, i, cat):
        """up`i` (1..4): transposed conv to level 4-i into the left columns of `cat`, two residual blocks."""
        lvl = 4 - i
        self.up[i - 1](y, m.nbr_up[lvl], m.n[lvl], out=cat[:, : self.up[i - 1].cout])
        r = self.dec[i - 1][0](cat, m.nbr3[lvl], m.n[lvl])
        return self.dec[i - 1][1](r, m.nbr3[lvl], m.n[lvl])

    def _cat_buffers(self, m):
        dev, dt = self.device, self.dtype
        # cat_i = [up_i output | encoder skip] at level 4-i
        widths = [self.up[i].cout + self.dec[i][0].c1.cin - self.up[i].cout for i in range(4)]
        return [_alloc(m.n[3 - i], widths[i], dt, dev) for i in range(4)]

    @torch.no_grad()
    def __call__(self, coords, feats, return_feat: bool = False):
        """logits f32 [N, n_cls]; with ``return_feat`` also the model's second output (network/minkunet.py:122 ``y4.F``,
        network/spvcnn.py:155 ``z3.F``): the 96-channel features in the engine's 16-bit activation type."""
        return self.forward(self.prepare(coords, feats), return_feat)

    @torch.no_grad()
    def prepare(self, coords, feats) -> "Prepared":
        """Everything that depends only on the coordinates: voxel grouping (SPVCNN), the 9 kernel maps and the point <-> voxel
        queries.  This is the only phase with host round trips (row counts of the levels), so callers that stream batches run
        it on a side stream while the previous batch's ``forward`` occupies the GPU (``StreamPipeline``)."""
        L.require_cuda(coords, feats)
        pr = Prepared()
        coords = coords.contiguous()
        if self.is_spvcnn:
            pr.zc, vcoords, pr.feats, level_counts = self._initial_voxelize(coords, feats)
            m = pr.m = self._maps(vcoords, level_counts)
            pr.q = {}
            for lvl in (0, 4, 2):
                pr.q[lvl] = self._corner_query(pr.zc, m, lvl) + self._cell_query(pr.zc, m, lvl)
        else:
            pr.m, pr.feats = self._maps(coords), feats
        return pr

    @torch.no_grad()
    def forward(self, pr: "Prepared", return_feat: bool = False):
        """The network proper on prepared maps: no host synchronisation, every launch is queued on the current stream."""
        logits, feat = self._spvcnn(pr) if self.is_spvcnn else self._minkunet(pr)
        return (logits, feat) if return_feat else logits

    @staticmethod
    def _maps(coords, level_counts=None):
        """All 9 kernel maps of a batch.  SURVEY 8d per map: 16 * N_ref + 16 * N_out + 4 * K * N_out (+ tables 24 * N_ref)."""
        box = {}
        with _stage("map_build", lambda: box["m"].algorithmic_bytes()):
            box["m"] = Maps(coords, level_counts)
        return box["m"]

    def _minkunet(self, pr):
        m, feats = pr.m, pr.feats
        cats = self._cat_buffers(m)
        x = self._stem(self._pad8(feats), m, cats[3])
        for lvl in range(1, 5):
            skip_out = cats[3 - lvl][:, self.up[3 - lvl].cout:] if lvl < 4 else None
            x = self._encode(x, m, lvl, out=skip_out)
        y = x
        for i in range(1, 5):
            y = self._decode(y, m, i, cats[i - 1])
        logits = self.classifier(y, None, m.n[0], out_dtype=torch.float32)
        return logits[:, : self.n_cls], y

    # ---- SPVCNN point branch
    def _table_of(self, m, lvl):
        return m.tables[lvl]

    def _corner_query(self, pts, m, lvl):
        n = pts.shape[0]
        idx = torch.empty((n, 8), dtype=torch.int, device=pts.device)
        w = torch.empty((n, 8), dtype=torch.float32, device=pts.device)
        t = m.tables[lvl]
        with _stage("point_query", n * (16 + 64) + 12 * m.n[lvl]):
            L.check(L.lib().lb_point_corner_query(L.ptr(pts), pts.stride(0), n, 2 ** lvl, L.ptr(t[0]), t[1], L.ptr(idx),
                                                  L.ptr(w), L.stream()))
        return idx, w

    def _cell_query(self, pts, m, lvl, segments=True):
        """network/utils.py:42-49 fused: point -> voxel index and its histogram.  ``segments`` (the engine's use) returns the
        atomic-free form instead: (seg_ptr, order) of lb_segment_order."""
        n = pts.shape[0]
        idx = torch.empty(n, dtype=torch.int, device=pts.device)
        t = m.tables[lvl]
        L.check(L.lib().lb_point_cell_query(L.ptr(pts), pts.stride(0), n, 2 ** lvl, L.ptr(t[0]), t[1], L.ptr(idx), L.stream()))
        counts = torch.empty(m.n[lvl], dtype=torch.int, device=pts.device)
        L.check(L.lib().lb_count(L.ptr(idx), n, L.ptr(counts), m.n[lvl], L.stream()))
        if not segments:
            return idx, counts
        # the points of every voxel, contiguous per voxel: voxelize becomes an atomic-free segmented mean (coordinates only,
        # so it is built in the prepare phase)
        seg_ptr = torch.empty(m.n[lvl] + 1, dtype=torch.int, device=pts.device)
        order = torch.empty(n, dtype=torch.int, device=pts.device)
        nb = L.lib().lb_segment_order_ws_bytes(m.n[lvl])
        ws = torch.empty(nb, dtype=torch.uint8, device=pts.device)
        L.check(L.lib().lb_segment_order(L.ptr(idx), n, L.ptr(counts), m.n[lvl], L.ptr(seg_ptr), L.ptr(order), L.ptr(ws), nb, L.stream()))
        return seg_ptr, order

    def _devox(self, x, idx, w, out_dtype=None):
        n, c = idx.shape[0], x.shape[1]
        out = torch.empty((n, c), dtype=out_dtype or self.dtype, device=x.device)
        # devoxelize = Np * (64 + C * e) + Nv * C * e
        with _stage("devoxelize", n * (64 + c * out.element_size()) + x.shape[0] * c * x.element_size()):
            L.check(L.lib().lb_devoxelize_fwd_ex(L.ptr(x), L.DT_OF[x.dtype], x.stride(0), L.ptr(idx), L.ptr(w), n, x.shape[0], c,
                                                 L.ptr(out), L.DT_OF[out.dtype], out.stride(0), L.stream()))
        return out

    def _vox(self, f, idx, counts, m_rows):
        acc = torch.empty((m_rows, f.shape[1]), dtype=torch.float32, device=f.device)
        # voxelize = (Np * e_in + Nv * 4) * C + 4 * Np
        with _stage("voxelize", (f.shape[0] * f.element_size() + m_rows * 4) * f.shape[1] + 4 * f.shape[0]):
            L.check(L.lib().lb_voxelize_fwd_ex(L.ptr(f), L.DT_OF[f.dtype], f.stride(0), L.ptr(idx), L.ptr(counts), f.shape[0],
                                               m_rows, f.shape[1], L.ptr(acc), L.stream()))
        return acc

    def _vox_seg(self, f, seg_ptr, order, m_rows):
        """point_to_voxel (network/utils.py:38-61) as a segmented mean: 16-bit in, 16-bit out, no atomics."""
        out = _alloc(m_rows, f.shape[1], f.dtype, f.device)
        # voxelize = (Np + Nv) * C * e + 4 * Np
        with _stage("voxelize", (f.shape[0] + m_rows) * f.shape[1] * f.element_size() + 4 * f.shape[0]):
            L.check(L.lib().lb_voxelize_segments(L.ptr(f), L.DT_OF[f.dtype], f.stride(0), L.ptr(order), L.ptr(seg_ptr), m_rows, f.shape[1],
                                                 L.ptr(out), out.stride(0), L.stream()))
        return out

    def _initial_voxelize(self, coords, feats):
        """network/utils.py:13-33: voxels = unique hashes of the floored point coordinates (the reference orders them by
        ascending hash; the engine keeps first-occurrence order, which only permutes internal rows)."""
        dev = coords.device
        zc = coords.float()
        zc = torch.cat([(zc[:, :3] * self.pres) / self.vres, zc[:, 3:]], 1).contiguous()       # new_float_coord
        cell = torch.floor(zc)
        h = F.sphash(cell.int())
        n = h.shape[0]
        # voxel order inside the engine is free (outputs are per point): group by hash in first-occurrence order
        cnt = _counters(dev)
        inv = torch.empty(n, dtype=torch.int, device=dev)
        first = torch.empty(n, dtype=torch.int, device=dev)
        nbytes = L.lib().lb_group_by_key_ws_bytes(n)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        L.check(L.lib().lb_group_by_key(L.ptr(h), n, L.ptr(inv), L.ptr(first), cnt.ptr(7), L.ptr(ws), nbytes, L.stream()))
        cell_i = cell.int()
        lc_ws = queue_level_counts(cell_i, cnt, 8)      # the whole pyramid's row counts ride on the same round trip as nv
        nv = cnt.read(7)
        level_counts = [int(cnt.buf[8 + l]) for l in range(4)]
        del lc_ws
        counts = torch.empty(nv, dtype=torch.int, device=dev)
        L.check(L.lib().lb_count(L.ptr(inv), n, L.ptr(counts), nv, L.stream()))
        vcoords = torch.empty((nv, 4), dtype=torch.int, device=dev)     # every member of a voxel has the same floored coords
        L.check(L.lib().lb_gather_rows16(L.ptr(cell_i), L.ptr(first), nv, L.ptr(vcoords), L.stream()))
        vfeats = self._vox(feats.contiguous(), inv, counts, nv)
        return zc, vcoords, vfeats, level_counts

    def _spvcnn(self, pr):
        m, vfeats = pr.m, pr.feats
        (iq0, w0, ci0, cn0), (iq4, w4, ci4, cn4), (iq2, w2, ci2, cn2) = pr.q[0], pr.q[4], pr.q[2]
        cats = self._cat_buffers(m)
        x0 = self._stem(self._pad8(vfeats), m, cats[3])
        z0 = self._devox(x0, iq0, w0)                                            # [Np, 32]
        x = self._vox_seg(z0, ci0, cn0, m.n[0])
        for lvl in range(1, 5):
            skip_out = cats[3 - lvl][:, self.up[3 - lvl].cout:] if lvl < 4 else None
            x = self._encode(x, m, lvl, out=skip_out)
        z1 = self.mlp[0](z0, None, z0.shape[0], residual=self._devox(x, iq4, w4), relu_first=True)   # [Np, 256]
        y = self._vox_seg(z1, ci4, cn4, m.n[4])                                  # dropout: identity in eval
        y = self._decode(y, m, 1, cats[0])
        y = self._decode(y, m, 2, cats[1])
        z2 = self.mlp[1](z1, None, z1.shape[0], residual=self._devox(y, iq2, w2), relu_first=True)   # [Np, 128]
        y = self._vox_seg(z2, ci2, cn2, m.n[2])
        y = self._decode(y, m, 3, cats[2])
        y = self._decode(y, m, 4, cats[3])
        z3 = self.mlp[2](z2, None, z2.shape[0], residual=self._devox(y, iq0, w0), relu_first=True)   # [Np, 96]
        logits = self.classifier(z3, None, z3.shape[0], out_dtype=torch.float32)
        return logits[:, : self.n_cls], z3


class Prepared:
    """Coordinate-only state of one batch (``InferenceEngine.prepare``): maps ``m``, input features ``feats`` (voxel means
    for SPVCNN), float point coordinates ``zc`` and the per-level point queries ``q[lvl] = (corner idx, weights, cell idx, counts)``."""
    __slots__ = ("m", "feats", "zc", "q", "ready")

    def __init__(self):
        self.m = self.feats = self.zc = self.ready = None
        self.q = {}

    def tensors(self):
        m = self.m
        for t in m.coords:
            yield t
        for tab in m.tables:
            yield tab[0]
        for group in (m.nbr3, m.nbr_dn, m.nbr_up):
            for item in group:
                for t in (item if isinstance(item, tuple) else (item,)):
                    yield t._base if t._base is not None else t
                if getattr(item, "tile_masks", None) is not None:
                    yield item.tile_masks
        for t in (self.feats, self.zc):
            if t is not None:
                yield t
        for q in self.q.values():
            yield from q


_ALLOC_CLASSES_SET = False


def _allocator_size_classes():
    """Consecutive batches differ by a few per cent in size; with exact-size blocks the caching allocator keeps splitting a
    cached block for the smaller request and calling cudaMalloc for the next larger one, and a cudaMalloc on a busy GPU blocks
    the launch loop for 10-110 ms (profiles/r02_host_stalls.txt).  Size classes of 1/8 octave make the blocks interchangeable.
    Respected if the user configured the allocator already (PYTORCH_CUDA_ALLOC_CONF)."""
    global _ALLOC_CLASSES_SET
    if _ALLOC_CLASSES_SET or _os.environ.get("PYTORCH_CUDA_ALLOC_CONF"):
        return
    _ALLOC_CLASSES_SET = True
    try:
        torch.cuda.memory._set_allocator_settings("roundup_power2_divisions:8")
    except Exception:          # noqa: BLE001 -- older / different allocator back ends: keep their behaviour
        pass


class StreamPipeline:
    """Streams batches through the engine with the host round trips off the critical path: ``prepare`` of batch i runs on a
    side stream (its row-count read-backs block the host only), ``forward`` of batch i on the caller's stream.  Because
    ``submit(i)`` returns as soon as forward(i) is QUEUED, the next ``submit`` prepares batch i+1 while the GPU is still
    busy with forward(i): map construction and its synchronisation bubbles hide behind the convolutions."""

    def __init__(self, engine: "InferenceEngine"):
        import collections
        self.engine = engine
        # High priority: the side stream's kernels are small and latency-bound, and the host blocks on their row counts five
        # times per batch.  At default priority they queue behind the thread blocks of the previous batch's convolutions and
        # prepare() took 6.5 ms of host time per step -- longer than the network itself, so the main stream starved.
        self.prep_stream = torch.cuda.Stream(device=engine.device, priority=PREP_PRIORITY)
        self._live = collections.deque()
        self.last_host_ms = (0.0, 0.0, 0.0)
        self._reserved = False
        self._reserved_main = False
        _allocator_size_classes()



    def prepare(self, coords, feats, after=None, wait_main=True) -> Prepared:
        """``after``: event that makes the inputs valid (e.g. their H2D copy); else ``wait_main`` orders the side stream behind
        everything queued on the caller's stream -- pass False when the inputs are already complete (resident batches, or
        tensors produced on the side stream itself), otherwise the previous batch's forward would be waited for."""
        main = torch.cuda.current_stream(self.engine.device)
        with torch.cuda.stream(self.prep_stream):
            if after is not None:
                self.prep_stream.wait_event(after)          # e.g. the H2D copy of this batch
            elif wait_main:
                self.prep_stream.wait_stream(main)          # inputs produced on the caller's stream
            if not self._reserved:
                # First batch: learn what one prepare() allocates on the side stream and park a segment of RESERVE_FACTOR (6) times
                # that in the stream's pool.  Later batches carve their buffers out of it, so no cudaMalloc (10-600 ms on a busy
                # GPU, profiles/r02_host_stalls.txt) can land in a steady-state step while the pool is still growing.
                self._reserved = True
                dev = self.engine.device
                before = torch.cuda.memory_allocated(dev)
                pr = self.engine.prepare(coords, feats)
                grown = torch.cuda.memory_allocated(dev) - before
                if grown > 0 and RESERVE_FACTOR > 0:
                    parked = torch.empty(int(grown * RESERVE_FACTOR), dtype=torch.uint8, device=dev)
                    del parked
            else:
                pr = self.engine.prepare(coords, feats)
            pr.ready = torch.cuda.Event()
            pr.ready.record(self.prep_stream)
        return pr

    def retire(self, *objects):
        """Keep ``objects`` (tensors allocated on the side stream but read by kernels queued on the caller's stream) alive until
        that queued work has finished, then let them go.  Holding references -- instead of ``Tensor.record_stream`` -- keeps
        the caching allocator's per-stream pools in steady state: nothing is ever deferred, nothing forces a fresh cudaMalloc."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.engine.device))
        self._live.append((objects, ev))
        while self._live and self._live[0][1].query():
            self._live.popleft()

    def submit(self, coords, feats, return_feat=False, after=None, wait_main=True):
        t0 = _time.perf_counter()
        pr = self.prepare(coords, feats, after, wait_main)
        t1 = _time.perf_counter()
        torch.cuda.current_stream(self.engine.device).wait_event(pr.ready)
        if not self._reserved_main and RESERVE_FACTOR > 0:
            # same for the caller's stream: park RESERVE_FACTOR / 2 times the peak of one forward() in its pool (activations of consecutive batches
            # differ by a few per cent in size; the one cudaMalloc that used to follow a few steps later cost 10-50 ms)
            self._reserved_main = True
            dev = self.engine.device
            torch.cuda.reset_peak_memory_stats(dev)
            before = torch.cuda.memory_allocated(dev)
            out = self.engine.forward(pr, return_feat)
            peak = torch.cuda.max_memory_allocated(dev) - before
            if peak > 0:
                parked = torch.empty(int(peak * RESERVE_FACTOR / 2), dtype=torch.uint8, device=dev)
                del parked
        else:
            out = self.engine.forward(pr, return_feat)
        t2 = _time.perf_counter()
        self.retire(pr)
        self.last_host_ms = ((t1 - t0) * 1e3, (t2 - t1) * 1e3, (_time.perf_counter() - t2) * 1e3)   # prepare, forward, retire
        return out


class HostPipeline:
    """Host-buffer front end of the engine: ``submit(coords_host, feats_host)`` / ``collect()``.

    The reference's loop (score/prob_inference.py:94-100) does ``.cuda()`` -> model -> ``.cpu()`` serially.  Here the H2D
    copy of batch i+1 and the D2H copy of logits i-1 run on their own streams (one per direction, so an upload never
    queues behind a download that is still waiting for its kernels) while batch i computes: PCIe traffic hides behind
    the kernels.  Inputs must be pinned for the copies to be asynchronous; outputs land in pinned buffers.
    """

    def __init__(self, engine: "InferenceEngine", depth: int = 2):
        self.engine = engine
        self.h2d_stream = torch.cuda.Stream(device=engine.device)
        self.d2h_stream = torch.cuda.Stream(device=engine.device)
        self.depth = depth
        self.stream_pipe = StreamPipeline(engine)
        self.pending = []          # (logits_dev, out_host, done_event)
        self._out_pool = {}
        self._in_pool = {}
        self._n_submitted = 0
        self.last_host_ms = (0.0, 0.0, 0.0, 0.0)

    def _out_buffer(self, shape, slot):
        key = (slot, shape[1])
        buf = self._out_pool.get(key)
        if buf is None or buf.shape[0] < shape[0]:
            buf = torch.empty((int(shape[0] * 1.25) + 1, shape[1]), dtype=torch.float32).pin_memory()
            self._out_pool[key] = buf
        return buf[: shape[0]]

    def _in_buffers(self, slot, coords_host, feats_host):
        """Persistent device input buffers per ring slot (no caching-allocator traffic on the copy streams)."""
        n = coords_host.shape[0]
        ent = self._in_pool.get(slot)
        if ent is None or ent[0].shape[0] < n or ent[1].shape[1] != feats_host.shape[1] or ent[0].dtype != coords_host.dtype \
                or ent[1].dtype != feats_host.dtype:
            cap = int(n * 1.25) + 1
            dev = self.engine.device
            ent = [torch.empty((cap, coords_host.shape[1]), dtype=coords_host.dtype, device=dev),
                   torch.empty((cap, feats_host.shape[1]), dtype=feats_host.dtype, device=dev), None]
            self._in_pool[slot] = ent
            # the caching allocator may hand back a block that kernels still queued on the compute stream are reading
            # (activations of the previous step are released as soon as engine() returns): the copy stream must not
            # write into it before they have finished
            self.h2d_stream.wait_stream(torch.cuda.current_stream(dev))
        return ent

    def submit(self, coords_host: torch.Tensor, feats_host: torch.Tensor):
        """Queue one batch (host tensors).  Returns immediately; results come back in order from ``collect``."""
        dev = self.engine.device
        t_a = _time.perf_counter()
        slot = self._n_submitted % (self.depth + 2)
        ent = self._in_buffers(slot, coords_host, feats_host)
        n = coords_host.shape[0]
        with torch.cuda.stream(self.h2d_stream):
            if ent[2] is not None:
                self.h2d_stream.wait_event(ent[2])        # the step that last read this slot has finished
            c = ent[0][:n]
            f = ent[1][:n]
            c.copy_(coords_host, non_blocking=True)
            f.copy_(feats_host, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self.h2d_stream)
        cur = torch.cuda.current_stream(dev)
        cur.wait_event(ready)
        t_b = _time.perf_counter()
        logits = self.stream_pipe.submit(c, f, after=ready)   # maps of this batch are built while the previous forward runs
        t_c = _time.perf_counter()
        computed = torch.cuda.Event()
        computed.record(cur)
        ent[2] = computed
        out = self._out_buffer(logits.shape, slot)        # ring: never a buffer still in flight
        self._n_submitted += 1
        with torch.cuda.stream(self.d2h_stream):
            self.d2h_stream.wait_event(computed)
            out.copy_(logits, non_blocking=True)
            logits.record_stream(self.d2h_stream)
            done = torch.cuda.Event()
            done.record(self.d2h_stream)
        self.pending.append((out, done))
        t_d = _time.perf_counter()
        results = []
        while len(self.pending) > self.depth:
            results.append(self._pop())
        self.last_host_ms = ((t_b - t_a) * 1e3, (t_c - t_b) * 1e3, (t_d - t_c) * 1e3, (_time.perf_counter() - t_d) * 1e3)   # upload, engine, download queue, wait for a result
        return results

    def _pop(self):
        out, done = self.pending.pop(0)
        done.synchronize()
        return out

    def collect(self):
        """Wait for and return all outstanding results (host float32 [N, n_cls] tensors, pinned; valid until reused)."""
        results = []
        while self.pending:
            results.append(self._pop())
        return results


# ------------------------------------------------------------------------------------------- zero-edit drop-in
def accelerate(model, dtype=torch.bfloat16):
    """Route ``model``'s eval-mode forward through ``InferenceEngine`` WITHOUT changing the object the caller holds:

        model = accelerate(model)            # after load_state_dict / .cuda() / .eval(); before or after the DDP wrap
        logits, out_feat = model(SparseTensor(feats, coords))          # score/prob_inference.py:97, evaluate.py:102 unchanged

    The instance keeps its class, parameters and ``state_dict`` keys (only the bound ``forward`` is replaced), so the
    reference's unmodified ``score/prob_inference.py`` / ``evaluate.py`` get the fused path with this one added line.
    Under ``torch.no_grad()`` in eval mode the call returns what the reference's forward returns (network/minkunet.py:122,
    network/spvcnn.py:155): logits f32 [N, n_cls] and the 96-channel features as f32.  In training mode, or with autograd
    enabled, the original module-by-module forward (the compat layer, autograd-correct) runs instead.
    The engine folds BatchNorm and packs weights when it is built; it is rebuilt when any parameter or buffer was replaced or
    written in place (storage pointer / ``_version``), which covers ``load_state_dict``, ``.to()``, optimiser steps.  Writes
    through ``.data`` are invisible to that check: call ``model.lidal_engine_invalidate()`` after such surgery."""
    import types
    inner = model.module if hasattr(model, "module") and not hasattr(model, "stem") else model     # DataParallel / DDP wrapper
    if getattr(inner, "_lidal_accelerated", False):
        return model
    original = inner.forward                                   # bound method of the reference class
    state = {"engine": None, "sig": None}

    def signature():
        return tuple((t.data_ptr(), t._version) for t in list(inner.parameters()) + list(inner.buffers()))

    def forward(self, x):
        if self.training or torch.is_grad_enabled():
            return original(x)
        sig = signature()
        if state["engine"] is None or sig != state["sig"]:
            state["engine"], state["sig"] = InferenceEngine(self, dtype), sig
        logits, feat = state["engine"](x.C, x.F, return_feat=True)
        return logits, feat.float()

    def invalidate(self):
        state["engine"] = None

    inner.forward = types.MethodType(forward, inner)
    inner.lidal_engine_invalidate = types.MethodType(invalidate, inner)
    inner._lidal_accelerated = True
    return model
