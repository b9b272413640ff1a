"""Device-side LiDAL scoring behind the reference's entry points (boundary B2).

``tta_tail``                     score/prob_inference.py:100-113 on device
``SequenceScorer``               device-resident frames of one sequence (prob + registered xyz + regions)
``init_worker`` / ``worker_func`` same names / argument meaning / return tuple as score/sv_level/LiDAL.py:17-103;
                                 the file lists may be the reference's .npy / .pickle paths (read once, then
                                 resident in HBM) or in-memory arrays.
``select_regions``               score/sv_level/LiDAL.py:230-325: device radix sort + device 5 m neighbour lists,
                                 native host replay of the greedy walk (``lb_select_walk``) with the iteration order
                                 of a CPython ``set`` modelled slot-exactly and checked against the interpreter (bit-exact ids).
"""
from __future__ import annotations

import ctypes as C
import pickle

import numpy as np
import torch

from .. import _lib as L

import os as _os
GRID_CELL_FACTOR = float(_os.environ.get('LIDAL_CELL_FACTOR', 1.05))         # grid cell just above dis_thresh (measured best on B200: near-sensor LiDAR
                               # rings are dense, so small cells minimise fp64 candidate evaluations); the 27-cell neighbourhood
                               # covers the match radius and cells whose box is farther than the radius are pruned exactly


def _ws(nbytes, device):
    return torch.empty(int(nbytes), dtype=torch.uint8, device=device)


# ------------------------------------------------------------------------------------------- prob_inference tail
def tta_tail(logits: torch.Tensor, inverse_indices: torch.Tensor, inf_reps: int, out_feat: torch.Tensor | None = None):
    """logits f32 [Nv, C] (all views stacked), inverse_indices int64 [inf_reps * Np]
    -> (prob_map_mean f32 [Np, C], pred int64 [Np]), both on device.  With ``out_feat`` (the model's second output,
    [Nv, 96], any float type) also its per-point mean over the views, f32 [Np, 96] (score/prob_inference.py:103-105,116-118)."""
    L.require_cuda(logits, inverse_indices, out_feat)
    logits = logits.contiguous().float()
    inv = inverse_indices.contiguous().long()
    assert inv.numel() % inf_reps == 0
    n_pts, n_cls = inv.numel() // inf_reps, logits.shape[1]
    prob = torch.empty((n_pts, n_cls), dtype=torch.float32, device=logits.device)
    pred = torch.empty(n_pts, dtype=torch.int64, device=logits.device)
    L.check(L.lib().lb_tta_softmax_mean_argmax(L.ptr(logits), logits.shape[0], n_cls, L.ptr(inv), inf_reps, n_pts,
                                               L.ptr(prob), L.ptr(pred), L.stream()))
    if out_feat is None:
        return prob, pred
    assert out_feat.shape[0] == logits.shape[0] and out_feat.stride(1) == 1
    c = out_feat.shape[1]
    feat = torch.empty((n_pts, c), dtype=torch.float32, device=logits.device)
    L.check(L.lib().lb_tta_feat_mean(L.ptr(out_feat), L.DT_OF[out_feat.dtype], out_feat.stride(0), out_feat.shape[0], L.ptr(inv),
                                     inf_reps, n_pts, c, L.ptr(feat), L.stream()))
    return prob, pred, feat


def register_points(raw: torch.Tensor, pose) -> torch.Tensor:
    """dataset/prepare_kdtree_sk.py:76-80 on device: raw f32 [Np, >=3] sensor-frame points, pose 4x4 float64 (sensor -> world)
    -> registered float64 [Np, 3], bit-equal to the numpy expression."""
    L.require_cuda(raw)
    raw = raw.float()
    if raw.stride(1) != 1:
        raw = raw.contiguous()
    n = raw.shape[0]
    xyz = torch.empty((n, 3), dtype=torch.float64, device=raw.device)
    pose = np.ascontiguousarray(np.asarray(pose, dtype=np.float64)).reshape(16)
    L.check(L.lib().lb_register_points(L.ptr(raw), raw.stride(0), n, (C.c_double * 16)(*pose.tolist()), L.ptr(xyz), L.stream()))
    return xyz


def regions_to_csr(sv2point, device):
    """The reference's ragged ``sv2point`` lists -> (region_ptr int32 [R+1], region_pts int32 [sum]) on device."""
    ptr = np.zeros(len(sv2point) + 1, np.int32)
    ptr[1:] = np.cumsum([len(p) for p in sv2point])
    pts = np.concatenate([np.asarray(p, np.int32) for p in sv2point]) if len(sv2point) else np.zeros(0, np.int32)
    return torch.from_numpy(ptr).to(device), torch.from_numpy(pts.astype(np.int32)).to(device)


# ------------------------------------------------------------------------------------------- per-frame scoring
def neighbour_ids(fid: int, n_frames: int, nei_num: int = 24):
    """LiDAL.py:41-42 (12 before + 12 after, reflected at the sequence ends)."""
    half = nei_num // 2
    before = [fid - o - 1 if fid - o - 1 >= 0 else half + o + 1 for o in range(half)]
    after = [fid + o + 1 if fid + o + 1 <= n_frames - 1 else n_frames - 2 - half - o for o in range(half)]
    return before + after


class Frame:
    __slots__ = ("xyz", "prob", "grid", "grid_bytes", "n", "region_ptr", "region_pts", "sv_id")


class SequenceScorer:
    """Frames of one sequence resident in HBM: registered xyz f64 [Np,3], prob f32 [Np,C], one hash grid each."""

    def __init__(self, device="cuda", nei_num=24, dis_thresh=0.1, n_total=None):
        self.device = torch.device(device)
        self.nei_num, self.dis_thresh = nei_num, dis_thresh
        self.cell = dis_thresh * GRID_CELL_FACTOR
        self.frames: dict[int, Frame] = {}          # frame id -> resident frame (a rank may hold only own + halo frames)
        self.n_total = n_total                      # frames in the whole sequence (default: all are resident)

    @property
    def n_frames(self):
        return self.n_total if self.n_total is not None else len(self.frames)

    def add_frame(self, xyz, prob, sv_id=None, sv2point=None, fid=None):
        f = Frame()
        f.xyz = torch.as_tensor(np.ascontiguousarray(xyz, dtype=np.float64) if isinstance(xyz, np.ndarray) else xyz
                                ).to(self.device, torch.float64).contiguous()
        f.prob = torch.as_tensor(prob).to(self.device, torch.float32).contiguous()
        f.n = f.xyz.shape[0]
        assert f.prob.shape[0] == f.n
        f.grid_bytes = L.lib().lb_frame_grid_bytes(f.n)
        f.grid = _ws(f.grid_bytes, self.device)
        ws_bytes = L.lib().lb_frame_grid_ws_bytes(f.n)
        ws = _ws(ws_bytes, self.device)
        L.check(L.lib().lb_frame_grid_build(L.ptr(f.xyz), f.n, self.cell, L.ptr(f.grid), f.grid_bytes, L.ptr(ws), ws_bytes,
                                            L.stream()))
        f.sv_id = None
        if sv2point is not None:
            self.set_regions(f, sv_id, sv2point)
        self.frames[len(self.frames) if fid is None else int(fid)] = f
        return f

    def set_regions(self, f: Frame, sv_id, sv2point):
        """sv2point: the reference's list of index arrays, or an already built device CSR pair (region_ptr, region_pts)."""
        if isinstance(sv2point, tuple) and len(sv2point) == 2 and torch.is_tensor(sv2point[0]):
            f.region_ptr, f.region_pts = (t.to(self.device, torch.int32).contiguous() for t in sv2point)
        else:
            f.region_ptr, f.region_pts = regions_to_csr(sv2point, self.device)
        f.sv_id = np.asarray(sv_id)

    def score_points(self, fid: int, want_nn=False):
        """LiDAL.py:59-81 for frame ``fid``: (interd f64 [Np], intere f32 [Np], matches int32 [Np][, nn int32 [Np,24]])."""
        q = self.frames[fid]
        nids = neighbour_ids(fid, self.n_frames, self.nei_num)
        refs = (L.FrameRef * len(nids))()
        for j, n in enumerate(nids):
            fr = self.frames[n]
            refs[j].grid, refs[j].xyz, refs[j].prob, refs[j].n = fr.grid.data_ptr(), fr.xyz.data_ptr(), fr.prob.data_ptr(), fr.n
        interd = torch.empty(q.n, dtype=torch.float64, device=self.device)
        intere = torch.empty(q.n, dtype=torch.float32, device=self.device)
        count = torch.empty(q.n, dtype=torch.int32, device=self.device)
        nn = torch.empty((q.n, len(nids)), dtype=torch.int32, device=self.device) if want_nn else None
        nb = L.lib().lb_interframe_score_ws_bytes(q.n, len(nids))
        ws = _ws(nb, self.device)
        L.check(L.lib().lb_interframe_score(L.ptr(q.grid), L.ptr(q.prob), q.n, q.prob.shape[1], refs, len(nids),
                                            self.dis_thresh, self.cell, L.ptr(interd), L.ptr(intere), L.ptr(count),
                                            L.ptr(nn), L.ptr(ws), nb, L.stream()))
        return (interd, intere, count, nn) if want_nn else (interd, intere, count)

    def score_frame_device(self, fid: int):
        """Per-region means on device: (sv_interds f32 [R], sv_interes f32 [R], sv_pnums i64 [R], sv_centers f32 [R,3])."""
        q = self.frames[fid]
        interd, intere, _ = self.score_points(fid)
        r = q.region_ptr.numel() - 1
        sv_d = torch.empty(r, dtype=torch.float32, device=self.device)
        sv_e = torch.empty(r, dtype=torch.float32, device=self.device)
        sv_n = torch.empty(r, dtype=torch.int64, device=self.device)
        sv_c = torch.empty((r, 3), dtype=torch.float32, device=self.device)
        L.check(L.lib().lb_region_reduce(L.ptr(interd), L.ptr(intere), L.ptr(q.xyz), L.ptr(q.region_ptr),
                                         L.ptr(q.region_pts), r, L.ptr(sv_d), L.ptr(sv_e), L.ptr(sv_n), L.ptr(sv_c),
                                         L.stream()))
        return sv_d, sv_e, sv_n, sv_c

    def score_frames_device(self, fids):
        """``score_frame_device`` for several frames; results concatenated on device in ``fids`` order, plus the matching
        sv_id vector (host int64).  No host synchronisation."""
        outs = [self.score_frame_device(f) for f in fids]
        cat = lambda i, shape, dt: (torch.cat([o[i] for o in outs]) if outs else torch.zeros(shape, dtype=dt, device=self.device))  # noqa: E731
        ids = np.concatenate([np.asarray(self.frames[f].sv_id, np.int64) for f in fids]) if len(fids) else np.zeros(0, np.int64)
        return ids, cat(0, 0, torch.float32), cat(1, 0, torch.float32), cat(2, 0, torch.int64), cat(3, (0, 3), torch.float32)

    def score_frame(self, fid: int, sv_pre=False):
        """The reference's ``worker_func`` return tuple (numpy, dtypes int64 / f32 / f32 / int64 / f32)."""
        sv_d, sv_e, sv_n, sv_c = self.score_frame_device(fid)
        out = (self.frames[fid].sv_id, sv_d.cpu().numpy(), sv_e.cpu().numpy())
        return out if sv_pre else out + (sv_n.cpu().numpy().astype(int), sv_c.cpu().numpy())


# ---- reference-shaped entry points (score/sv_level/LiDAL.py:17-103) ----
var_dict: dict = {}


def _load_xyz(src):
    if isinstance(src, str):
        with open(src, "rb") as f:
            tree = pickle.load(f)           # sklearn KDTree pickle written by dataset/prepare_kdtree_sk.py:83-88
        return np.asarray(tree.data)
    return np.asarray(src.data) if hasattr(src, "data") and not isinstance(src, np.ndarray) else np.asarray(src)


def init_worker(sv_pre, nei_num, dis_thresh, seq_id, prob_files, kdtree_files, sv_info_files, device="cuda"):
    """Same arguments as the reference; entries may be file paths (reference formats) or in-memory arrays."""
    scorer = SequenceScorer(device, nei_num, dis_thresh)
    for p, k, s in zip(prob_files, kdtree_files, sv_info_files):
        prob = np.load(p) if isinstance(p, str) else p
        if isinstance(s, str):
            with open(s, "rb") as f:
                sv_id, sv2point = pickle.load(f)
        else:
            sv_id, sv2point = s
        scorer.add_frame(_load_xyz(k), prob, sv_id, sv2point)
    var_dict.update(sv_pre=sv_pre, nei_num=nei_num, dis_thresh=dis_thresh, seq_id=seq_id, scorer=scorer)


def worker_func(id):
    return var_dict["scorer"].score_frame(int(id), var_dict["sv_pre"])


# ------------------------------------------------------------------------------------------- frame-level baselines
def frame_level_scores(prob: torch.Tensor):
    """(softmax-entropy, margin, least-confidence) frame scores of one prob map f32 [Np, C] on device
    (score/frame_level/softmax_entropy.py:34, margin_sampling.py:33-34, least_confidence_sampling.py)."""
    L.require_cuda(prob)
    prob = prob.contiguous().float()
    out = torch.empty(3, dtype=torch.float64, device=prob.device)
    nbytes = L.lib().lb_frame_level_ws_bytes()
    ws = _ws(nbytes, prob.device)
    L.check(L.lib().lb_frame_level_scores(L.ptr(prob), prob.shape[0], prob.shape[1], L.ptr(out), L.ptr(ws), nbytes, L.stream()))
    ent, mar, conf = out.cpu().tolist()
    return ent, mar, conf


def segment_entropy(pred: torch.Tensor, sv2point, n_cls: int) -> float:
    """score/frame_level/segment_entropy.py:41-48 for one frame: pred int64 [Np] on device, regions as the reference's
    ``sv2point`` lists or a device CSR pair."""
    L.require_cuda(pred)
    pred = pred.contiguous().long()
    ptr, pts = sv2point if (isinstance(sv2point, tuple) and torch.is_tensor(sv2point[0])) else regions_to_csr(sv2point, pred.device)
    r = ptr.numel() - 1
    out = torch.empty(1, dtype=torch.float64, device=pred.device)
    ws = _ws(max(r, 1) * 8, pred.device)
    L.check(L.lib().lb_segment_entropy(L.ptr(pred), pred.numel(), n_cls, L.ptr(ptr), L.ptr(pts), r, L.ptr(out), L.ptr(ws), ws.numel(),
                                       L.stream()))
    return float(out.item())


REDAL_ALPHA, REDAL_GAMMA, REDAL_FT_DIM = 1.0, 0.05, 96          # score/sv_level/ReDAL.py:15-21


def redal_region_scores(prob: torch.Tensor, outfeat: torch.Tensor, curvature, sv_id, sv2point, sv_pre: bool = False):
    """score/sv_level/ReDAL.py:37-84 (``worker_func``) for one frame on device: prob f32 [Np, C], outfeat f32 [Np, 96] (the
    TTA-mean out_feat of ``tta_tail``), curvature f32 [Np] or None.  Returns the reference's tuple
    (sv_id, sv_scores f32 [R], sv_feats f32 [R, 96][, sv_pnums int [R]]) as numpy."""
    L.require_cuda(prob, outfeat)
    dev = prob.device
    prob = prob.contiguous().float()
    outfeat = outfeat.contiguous().float()
    n, n_cls = prob.shape
    assert outfeat.shape[0] == n and outfeat.shape[1] == REDAL_FT_DIM
    curv = None if curvature is None else torch.as_tensor(curvature).to(dev, torch.float32).contiguous()
    ptr, pts = sv2point if (isinstance(sv2point, tuple) and torch.is_tensor(sv2point[0])) else regions_to_csr(sv2point, dev)
    r = ptr.numel() - 1
    point_score = torch.empty(n, dtype=torch.float32, device=dev)
    L.check(L.lib().lb_redal_point_scores(L.ptr(prob), n, n_cls, L.ptr(curv), REDAL_ALPHA, REDAL_GAMMA, L.ptr(point_score), L.stream()))
    sv_scores = torch.empty(r, dtype=torch.float32, device=dev)
    sv_feats = torch.empty((r, REDAL_FT_DIM), dtype=torch.float32, device=dev)
    L.check(L.lib().lb_region_mean_f32(L.ptr(point_score), L.ptr(ptr), L.ptr(pts), r, L.ptr(sv_scores), L.stream()))
    L.check(L.lib().lb_region_feat_mean(L.ptr(outfeat), outfeat.stride(0), REDAL_FT_DIM, L.ptr(ptr), L.ptr(pts), r, L.ptr(sv_feats),
                                        L.stream()))
    out = (np.asarray(sv_id), sv_scores.cpu().numpy(), sv_feats.cpu().numpy())
    if sv_pre:
        return out
    return out + ((ptr[1:] - ptr[:-1]).cpu().numpy().astype(int),)


# ------------------------------------------------------------------------------------------- multi-sequence driver
def score_dataset(sequences, nei_num=24, dis_thresh=0.1, n_regions_total=None, sv_pnums=None, sv_centers=None, device="cuda"):
    """score/sv_level/LiDAL.py:163-222: every sequence is scored on its own (no connections between sequences) and the
    per-region values are scattered into the global arrays by ``sv_id``.  ``sequences`` is a list of
    (prob_files, kdtree_files, sv_info_files) triples in ``train_split`` order -- file paths in the reference's formats or
    in-memory arrays, as for ``init_worker``.  If ``sv_pnums`` / ``sv_centers`` are given (the reference's cached
    sv_pnums.npy / sv_centers.npy, LiDAL.py:171-175) they are reused (``sv_pre``); otherwise they are computed, with
    ``idx * 1000.0`` added to the centres of sequence idx (LiDAL.py:218).
    Returns (sv_interds f32, sv_interes f32, sv_pnums int, sv_centers f32 [.,3], sv_pre)."""
    sv_pre = sv_pnums is not None and sv_centers is not None
    results = []
    for idx, (prob_files, kdtree_files, sv_info_files) in enumerate(sequences):
        assert len(prob_files) == len(kdtree_files) == len(sv_info_files)                          # :199-200
        init_worker(sv_pre, nei_num, dis_thresh, idx, prob_files, kdtree_files, sv_info_files, device=device)
        scorer = var_dict["scorer"]
        results.append((idx, scorer.score_frames_device(list(range(len(prob_files))))))
        scorer.frames.clear()                                                                      # free the sequence's HBM
    if n_regions_total is None:
        n_regions_total = 1 + max((int(ids.max()) for _, (ids, *_r) in results if len(ids)), default=-1)
    sv_interds = np.zeros(n_regions_total, np.float32)                                             # :164-166
    sv_interes = np.zeros(n_regions_total, np.float32)
    if not sv_pre:
        sv_pnums = np.zeros(n_regions_total, int)                                                  # :178-180
        sv_centers = np.zeros((n_regions_total, 3), np.float32)
    for idx, (ids, d, e, pn, c) in results:
        sv_interds[ids] = d.cpu().numpy()                                                          # :208-216
        sv_interes[ids] = e.cpu().numpy()
        if not sv_pre:
            sv_pnums[ids] = pn.cpu().numpy().astype(int)
            sv_centers[ids] = c.cpu().numpy() + idx * 1000.0                                       # :218 (float32 + python float -> float32)
    return sv_interds, sv_interes, sv_pnums, sv_centers, sv_pre


# ------------------------------------------------------------------------------------------- selection
def argsort_f32(keys: torch.Tensor) -> torch.Tensor:
    L.require_cuda(keys)
    keys = keys.contiguous().float()
    n = keys.numel()
    order = torch.empty(n, dtype=torch.int32, device=keys.device)
    nbytes = L.lib().lb_argsort_ws_bytes(n)
    ws = _ws(nbytes, keys.device)
    L.check(L.lib().lb_argsort_f32(L.ptr(keys), n, L.ptr(order), L.ptr(ws), nbytes, L.stream()))
    return order


def region_pairs(centers: torch.Tensor, radius: float, max_pairs=None):
    """CSR (row_ptr int64 [n+1], idx int32 [nnz]) of regions with float32 distance < radius (LiDAL.py:252-254): a hash grid
    over all centres, counted and filled on device.  Returns None when the lists would exceed ``max_pairs``."""
    L.require_cuda(centers)
    centers = centers.contiguous().float()
    n = centers.shape[0]
    counts = torch.zeros(n, dtype=torch.int32, device=centers.device)
    nbytes = L.lib().lb_region_pairs_ws_bytes(n)
    ws = _ws(nbytes, centers.device)
    L.check(L.lib().lb_region_pairs(L.ptr(centers), n, float(radius), L.ptr(counts), None, L.ptr(ws), nbytes, L.stream()))
    row_ptr = torch.zeros(n + 1, dtype=torch.int64, device=centers.device)
    row_ptr[1:] = torch.cumsum(counts, 0)
    nnz = int(row_ptr[-1].item())
    if max_pairs is not None and nnz > max_pairs:
        return None
    if nnz >= 2 ** 31:
        raise L.LidalError(f"region_pairs: {nnz} pairs exceed the int32 offsets of lb_region_pairs; build the lists in chunks")
    offs = row_ptr[:-1].to(torch.int32).contiguous()
    idx = torch.empty(max(nnz, 1), dtype=torch.int32, device=centers.device)
    L.check(L.lib().lb_region_pairs(L.ptr(centers), n, float(radius), L.ptr(offs), L.ptr(idx), L.ptr(ws), nbytes, L.stream()))
    return row_ptr, idx[:nnz]


_CELL_BIAS = 1 << 20
_CELL_NBRS = [(dx << 42) + (dy << 21) + dz for dx in (-1, 0, 1) for dy in (-1, 0, 1) for dz in (-1, 0, 1)]


def region_cells(centers: np.ndarray, radius: float):
    """One integer key per region: its cell in a grid of ``radius``-sized cells (anything closer than ``radius`` lies in
    the 27 surrounding cells).  Python ints so that dict lookups in the greedy walk stay cheap."""
    c = np.floor(np.asarray(centers, np.float64) / float(radius)).astype(np.int64) + _CELL_BIAS
    return ((c[:, 0] << 42) | (c[:, 1] << 21) | c[:, 2]).tolist()


class _GridIndex:
    """The regions added so far, indexed by 5 m cells (cost independent of the dataset size); distances are the
    reference's float32 expression ``np.sqrt(np.square(a - b).sum())`` (LiDAL.py:252)."""

    def __init__(self, centers, cells, radius):
        self.centers, self.cells, self.thr, self.grid = centers, cells, np.float32(radius), {}

    def near(self, sv):
        key = self.cells[sv]
        pool = []
        for dk in _CELL_NBRS:
            lst = self.grid.get(key + dk)
            if lst:
                pool.extend(lst)
        if not pool:
            return pool
        arr = np.asarray(pool, dtype=np.int64)
        dist = np.sqrt(np.square(self.centers[sv] - self.centers[arr]).sum(1))     # float32, same operation order
        return arr[dist < self.thr].tolist()

    def add(self, sv):
        self.grid.setdefault(self.cells[sv], []).append(sv)

    def remove(self, sv):
        self.grid[self.cells[sv]].remove(sv)


class _CsrIndex:
    """All in-range pairs precomputed on device (``region_pairs``): a candidate's lookups are set-membership tests only."""

    EAGER_PAIRS = 1 << 22            # up to here the whole pair array becomes one Python list up front (slicing it is cheaper
                                     # than a numpy slice + tolist per candidate); above, rows are converted on demand

    def __init__(self, row_ptr, nbr_idx):
        self.row_ptr, self.members = row_ptr, set()
        self.eager = len(nbr_idx) <= self.EAGER_PAIRS
        self.nbr_idx = nbr_idx.tolist() if self.eager and not isinstance(nbr_idx, list) else nbr_idx

    def near(self, sv):
        m = self.members
        if not m:
            return []
        row = self.nbr_idx[self.row_ptr[sv]:self.row_ptr[sv + 1]]
        return list(m.intersection(row if self.eager else row.tolist()))

    def add(self, sv):
        self.members.add(sv)

    def remove(self, sv):
        self.members.discard(sv)


class _SetOrder:
    """Slot positions of a CPython ``set`` of small non-negative ints, replayed operation by operation (open addressing,
    9 linear probes, perturb shift 5, growth at fill*5 >= mask*3 to the first power of two above used*4 -- Objects/setobject.c).
    The reference's greedy walk stops at the FIRST in-range member in the set's iteration order (LiDAL.py:246-252); with
    the slots known that member is ``min(near, key=slot)`` instead of a scan over every added region.  The model is
    verified against the running interpreter's own ``set`` before it is trusted (``_set_model``)."""

    __slots__ = ("keys", "mask", "fill", "used", "slot", "last_dummy")

    def __init__(self, last_dummy=True):
        self.keys, self.mask, self.fill, self.used, self.slot, self.last_dummy = [None] * 8, 7, 0, 0, {}, last_dummy

    def add(self, key):
        keys, mask = self.keys, self.mask
        i, perturb, free = key & mask, key, -1
        while True:
            for j in range(i, i + 10 if i + 9 <= mask else i + 1):
                k = keys[j]
                if k is None:
                    if free >= 0:
                        j = free
                    else:
                        self.fill += 1
                    keys[j] = key
                    self.slot[key] = j
                    self.used += 1
                    if free < 0 and self.fill * 5 >= mask * 3:
                        self._resize(self.used * 2 if self.used > 50000 else self.used * 4)
                    return
                if k == key:
                    return
                if k == -1 and (self.last_dummy or free < 0):
                    free = j
            perturb >>= 5
            i = (i * 5 + 1 + perturb) & mask

    def remove(self, key):
        self.keys[self.slot.pop(key)] = -1       # dummy entry: still counted in fill, reusable by a later insertion
        self.used -= 1

    def _resize(self, minused):
        size = 8
        while size <= minused:
            size <<= 1
        mask, new, slot = size - 1, [None] * size, self.slot
        for key in self.keys:                      # old table order, clean insertion (no dummies in the new table)
            if key is None or key == -1:
                continue
            i, perturb = key & mask, key
            while True:
                if new[i] is None:
                    break
                if i + 9 <= mask:
                    for j in range(i + 1, i + 10):
                        if new[j] is None:
                            break
                    else:
                        j = -1
                    if j >= 0:
                        i = j
                        break
                perturb >>= 5
                i = (i * 5 + 1 + perturb) & mask
            new[i] = key
            slot[key] = i
        self.keys, self.mask, self.fill = new, mask, self.used

    def __iter__(self):
        return (k for k in self.keys if k is not None and k != -1)

    def __len__(self):
        return self.used


_SET_MODEL = []


def _set_model():
    """The ``_SetOrder`` variant whose iteration order equals the running interpreter's ``set`` on a randomised add /
    remove sequence crossing several resizes (cached); None when neither does -- the walk then scans a real set."""
    if not _SET_MODEL:
        rng = np.random.RandomState(12345)
        found = None
        for last_dummy in (True, False):
            ok = True
            for trial in range(3):
                real, model, members = set(), _SetOrder(last_dummy), []
                for step in range(6000):
                    if members and rng.rand() < 0.35:
                        key = members.pop(rng.randint(len(members)))
                        real.remove(key)
                        model.remove(key)
                    else:
                        key = int(rng.randint(0, 40000 * (trial + 1)))
                        if key not in real:
                            members.append(key)
                        real.add(key)
                        model.add(key)
                    if step % 500 == 499 and list(real) != list(model):
                        ok = False
                        break
                if not ok or list(real) != list(model):
                    ok = False
                    break
            if ok:
                found = last_dummy
                break
        _SET_MODEL.append(found)
    return _SET_MODEL[0]


def _greedy_walk(order, cand_ids, interds, interes, pnums, index, flags, flag_value, point_limit, prefer_higher_entropy, skip_zero):
    """Host replay of LiDAL.py:242-270 / 293-325.  The reference scans a CPython ``set`` of the regions added so far and
    stops at the FIRST member within 5 m; which member that is depends on the set's iteration order, so the same add /
    remove sequence is fed to a slot-exact model of the interpreter's ``set`` (``_SetOrder``, checked against the real one;
    a real ``set`` is scanned if the check fails) and its order is consulted only when two or more members are in range.
    ``index`` answers "which added regions lie within 5 m of this candidate" (_GridIndex or _CsrIndex)."""
    model = _set_model()
    added = set() if model is None else _SetOrder(model)
    slot_of = None if model is None else added.slot.__getitem__
    cand_list = cand_ids.tolist()
    interds, interes, pnums = np.asarray(interds).tolist(), np.asarray(interes).tolist(), np.asarray(pnums).tolist()   # exact
    for idx in order.tolist():
        sv = cand_list[idx]                                  # hash(int) == hash(np.int64): same set order as the reference
        if skip_zero and interds[sv] == 0:
            continue
        near = index.near(sv)
        if near:
            hit = near[0]
            if len(near) > 1:
                if slot_of is not None:
                    hit = min(near, key=slot_of)             # first in-range member in the set's iteration order
                else:
                    hit = next(filter(set(near).__contains__, added))
            better = interes[hit] < interes[sv] if prefer_higher_entropy else interes[hit] > interes[sv]
            if better:
                flags[sv] = flag_value
                flags[hit] = 0
                added.add(sv)
                added.remove(hit)
                index.remove(hit)
                index.add(sv)
                point_limit = point_limit + pnums[hit] - pnums[sv]
            continue
        point_limit -= pnums[sv]
        if point_limit < 0:
            break
        flags[sv] = flag_value
        added.add(sv)
        index.add(sv)
    return flags


def native_walk(visit, interds, interes, pnums, row_ptr, nbr_idx, flags, flag_value, point_limit, prefer_higher_entropy,
                skip_zero, last_dummy=True):
    """``lb_select_walk`` on host arrays: the same walk as ``_greedy_walk`` over CSR neighbour lists; ``flags`` (int numpy
    array) is updated in place.  ``visit``: region ids in visiting order."""
    as_p = lambda a: a.ctypes.data_as(C.c_void_p)            # noqa: E731
    keep = [np.ascontiguousarray(visit, dtype=np.int64), np.ascontiguousarray(interds, dtype=np.float64),
            np.ascontiguousarray(interes, dtype=np.float64), np.ascontiguousarray(pnums, dtype=np.int64),
            np.ascontiguousarray(row_ptr, dtype=np.int64), np.ascontiguousarray(nbr_idx, dtype=np.int32),
            np.ascontiguousarray(flags, dtype=np.int64)]
    n_added = C.c_int64(0)
    L.check(L.lib().lb_select_walk(as_p(keep[0]), keep[0].size, as_p(keep[1]), as_p(keep[2]), as_p(keep[3]), as_p(keep[4]),
                                   as_p(keep[5]), flags.size, as_p(keep[6]), int(flag_value), int(point_limit),
                                   int(bool(prefer_higher_entropy)), int(bool(skip_zero)), int(bool(last_dummy)), C.byref(n_added)))
    flags[:] = keep[6]
    return int(n_added.value)


CSR_MAX_PAIRS = 1 << 27          # above this the all-pairs lists are not built (memory); the grid index is used instead


def select_regions(sv_flags, sv_interds, sv_interes, sv_pnums, sv_centers, train_point_num, sv_dis_thresh=5.0,
                   device="cuda", method="auto", timings=None):
    """LiDAL.py:230-325.  Inputs are the global per-region arrays (numpy); returns int flags (0 / 1 labelled / 2 pseudo).

    The two sorts run on device (radix argsort of the unlabelled regions' divergence, LiDAL.py:235,283) and so does the
    5 m neighbourhood search (``region_pairs``, a hash grid over all centres) unless the pair lists would be too large
    (``method="grid"``: only the added regions are indexed, on the host).  The greedy walk is inherently sequential: it
    runs on the host, natively (``lb_select_walk``) over the device-built lists when the set-order model matches the
    interpreter (``method="auto"``/``"native"``), else as the Python replay (``"csr"``/``"grid"``).  Ties: the reference's
    ``np.argsort`` is an unstable introsort, so the visiting order of EQUAL keys is only defined by numpy itself -- when
    the device sort sees equal keys it falls back to ``np.argsort`` for that pass, so the selected ids stay identical to
    the reference's in every case.  ``timings`` (dict, optional) receives the host milliseconds of each part."""
    import time as _time
    dev = torch.device(device)
    t0 = _time.perf_counter()
    flags = np.asarray(sv_flags).astype(int)
    interds, interes = np.asarray(sv_interds), np.asarray(sv_interes)
    pnums = np.asarray(sv_pnums)
    centers = np.ascontiguousarray(np.asarray(sv_centers, np.float32))
    d_dev = torch.as_tensor(np.asarray(sv_interds, np.float32)).to(dev)
    csr, native = None, None
    if method in ("auto", "csr", "native"):
        pairs = region_pairs(torch.from_numpy(centers).to(dev), sv_dis_thresh, None if method != "auto" else CSR_MAX_PAIRS)
        if pairs is not None:
            model = _set_model() if method in ("auto", "native") else None
            if method == "native" and model is None:
                raise L.LidalError("select_regions(method='native'): the set-order model does not match this interpreter")
            if model is not None:
                native = (np.ascontiguousarray(pairs[0].cpu().numpy()), np.ascontiguousarray(pairs[1].cpu().numpy()), model,
                          np.ascontiguousarray(interds, dtype=np.float64), np.ascontiguousarray(interes, dtype=np.float64),
                          np.ascontiguousarray(pnums, dtype=np.int64))
            else:
                csr = (pairs[0].cpu().tolist(), pairs[1].cpu().numpy())      # lists of a region are converted on demand
    cells = None if (csr is not None or native is not None) else region_cells(centers, sv_dis_thresh)
    t1 = _time.perf_counter()
    sort_s = [0.0]

    def sorted_candidates():
        h = _time.perf_counter()
        ids = np.where(flags == 0)[0]
        keys = d_dev[torch.from_numpy(ids).to(dev)]
        order_dev = argsort_f32(keys)
        sk = keys[order_dev.long()]
        has_ties = bool((sk[1:] == sk[:-1]).any().item()) if ids.size > 1 else False
        order = np.argsort(interds[ids]) if has_ties else order_dev.cpu().numpy()
        sort_s[0] += _time.perf_counter() - h
        return ids, order

    def walk(order, ids, flag_value, prefer_higher_entropy, skip_zero):
        if native is None:
            index = _CsrIndex(*csr) if csr is not None else _GridIndex(centers, cells, sv_dis_thresh)
            return _greedy_walk(order, ids, interds, interes, pnums, index, flags, flag_value, limit, prefer_higher_entropy, skip_zero)
        row_ptr, nbr_idx, model, d64, e64, pn64 = native
        native_walk(ids[order], d64, e64, pn64, row_ptr, nbr_idx, flags, flag_value, limit, prefer_higher_entropy, skip_zero, model)

    ids, order = sorted_candidates()                                           # :232-235
    limit = round(0.01 * train_point_num)                                      # :240
    walk(order[::-1], ids, 1, True, False)
    ids, order = sorted_candidates()                                           # :281-283 (before the reset)
    flags[flags == 2] = 0                                                      # :286
    walk(order, ids, 2, False, True)
    if timings is not None:
        t2 = _time.perf_counter()
        timings.update(select_pairs_ms=(t1 - t0) * 1e3, select_sort_ms=sort_s[0] * 1e3,
                       select_walk_ms=(t2 - t1 - sort_s[0]) * 1e3, select_path="native" if native is not None else
                       ("python-csr" if csr is not None else "python-grid"))
    return flags
