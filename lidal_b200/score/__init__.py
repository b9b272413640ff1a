"""Device-side LiDAL scoring behind the reference's entry points (boundary B2).

``tta_tail``                     score/prob_inference.py:100-113 on device
``SequenceScorer``               device-resident frames of one sequence (prob + registered xyz + regions)
``init_worker`` / ``worker_func`` same names / argument meaning / return tuple as score/sv_level/LiDAL.py:17-103;
                                 the file lists may be the reference's .npy / .pickle paths (read once, then
                                 resident in HBM) or in-memory arrays.
``select_regions``               score/sv_level/LiDAL.py:230-325: device radix sort + device 5 m neighbour lists,
                                 host replay of the greedy walk with a real CPython ``set`` (bit-exact ids).
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
def tta_tail(logits: torch.Tensor, inverse_indices: torch.Tensor, inf_reps: int):
    """logits f32 [Nv, C] (all views stacked), inverse_indices int64 [inf_reps * Np]
    -> (prob_map_mean f32 [Np, C], pred int64 [Np]), both on device."""
    L.require_cuda(logits, inverse_indices)
    logits = logits.contiguous().float()
    inv = inverse_indices.contiguous().long()
    assert inv.numel() % inf_reps == 0
    n_pts, n_cls = inv.numel() // inf_reps, logits.shape[1]
    prob = torch.empty((n_pts, n_cls), dtype=torch.float32, device=logits.device)
    pred = torch.empty(n_pts, dtype=torch.int64, device=logits.device)
    L.check(L.lib().lb_tta_softmax_mean_argmax(L.ptr(logits), logits.shape[0], n_cls, L.ptr(inv), inf_reps, n_pts,
                                               L.ptr(prob), L.ptr(pred), L.stream()))
    return prob, pred


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
        ptr = np.zeros(len(sv2point) + 1, np.int32)
        ptr[1:] = np.cumsum([len(p) for p in sv2point])
        pts = np.concatenate([np.asarray(p, np.int32) for p in sv2point]) if len(sv2point) else np.zeros(0, np.int32)
        f.region_ptr = torch.from_numpy(ptr).to(self.device)
        f.region_pts = torch.from_numpy(pts.astype(np.int32)).to(self.device)
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
        nn = torch.empty((q.n, len(nids)), dtype=torch.int32, device=self.device)
        L.check(L.lib().lb_interframe_score(L.ptr(q.grid), L.ptr(q.prob), q.n, q.prob.shape[1], refs, len(nids),
                                            self.dis_thresh, self.cell, L.ptr(interd), L.ptr(intere), L.ptr(count),
                                            L.ptr(nn), L.stream()))
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


def region_pairs(centers: torch.Tensor, radius: float):
    """CSR (row_ptr int64 [n+1], idx int32 [nnz]) of regions with float32 distance < radius (LiDAL.py:252-254)."""
    L.require_cuda(centers)
    centers = centers.contiguous().float()
    n = centers.shape[0]
    counts = torch.zeros(n, dtype=torch.int32, device=centers.device)
    nbytes = L.lib().lb_region_pairs_ws_bytes(n)
    ws = _ws(nbytes, centers.device)
    L.check(L.lib().lb_region_pairs(L.ptr(centers), n, float(radius), L.ptr(counts), None, L.ptr(ws), nbytes, L.stream()))
    row_ptr = torch.zeros(n + 1, dtype=torch.int64, device=centers.device)
    row_ptr[1:] = torch.cumsum(counts, 0)
    offs = row_ptr[:-1].to(torch.int32).contiguous()
    idx = torch.empty(max(int(row_ptr[-1].item()), 1), dtype=torch.int32, device=centers.device)
    L.check(L.lib().lb_region_pairs(L.ptr(centers), n, float(radius), L.ptr(offs), L.ptr(idx), L.ptr(ws), nbytes, L.stream()))
    return row_ptr, idx[: int(row_ptr[-1].item())]


def _greedy_walk(order, cand_ids, interds, interes, pnums, row_ptr, nbr_idx, flags, flag_value, point_limit,
                 prefer_higher_entropy, skip_zero):
    """Host replay of LiDAL.py:242-270 / 293-325.  The reference scans a CPython ``set`` and stops at the FIRST
    member within 5 m; which member that is depends on the set's iteration order, so the same add / remove
    sequence is fed to a real ``set`` and its order is consulted only when two or more members are in range."""
    added = set()
    for idx in order:
        if skip_zero and interds[cand_ids[idx]] == 0:
            continue
        sv = cand_ids[idx]                                   # np.int64, hashed like the reference's elements
        near = [j for j in nbr_idx[row_ptr[sv]:row_ptr[sv + 1]] if j in added]
        if near:
            hit = near[0]
            if len(near) > 1:
                inrange = set(near)
                hit = next(m for m in added if m in inrange)
            better = interes[hit] < interes[sv] if prefer_higher_entropy else interes[hit] > interes[sv]
            if better:
                flags[sv] = flag_value
                flags[hit] = 0
                added.add(sv)
                added.remove(hit)
                point_limit = point_limit + pnums[hit] - pnums[sv]
            continue
        point_limit -= pnums[sv]
        if point_limit < 0:
            break
        flags[sv] = flag_value
        added.add(sv)
    return flags


def select_regions(sv_flags, sv_interds, sv_interes, sv_pnums, sv_centers, train_point_num, sv_dis_thresh=5.0,
                   device="cuda"):
    """LiDAL.py:230-325.  Inputs are the global per-region arrays (numpy); returns int flags (0 / 1 labelled / 2 pseudo)."""
    dev = torch.device(device)
    flags = np.asarray(sv_flags).astype(int)
    d_dev = torch.as_tensor(np.asarray(sv_interds, np.float32)).to(dev)
    row_ptr, nbr_idx = region_pairs(torch.as_tensor(np.asarray(sv_centers, np.float32)).to(dev), sv_dis_thresh)
    row_ptr, nbr_idx = row_ptr.cpu().numpy(), nbr_idx.cpu().numpy().astype(np.int64)
    interds, interes = np.asarray(sv_interds), np.asarray(sv_interes)
    pnums = np.asarray(sv_pnums)

    def sorted_candidates():
        ids = np.where(flags == 0)[0]
        ids_dev = torch.from_numpy(ids).to(dev)
        order = argsort_f32(d_dev[ids_dev]).cpu().numpy()
        return ids, order

    ids, order = sorted_candidates()                                           # :232-235
    limit = round(0.01 * train_point_num)                                      # :240
    _greedy_walk(order[::-1], ids, interds, interes, pnums, row_ptr, nbr_idx, flags, 1, limit, True, False)
    ids, order = sorted_candidates()                                           # :281-283 (before the reset)
    flags[flags == 2] = 0                                                      # :286
    _greedy_walk(order, ids, interds, interes, pnums, row_ptr, nbr_idx, flags, 2, limit, False, True)
    return flags
