"""Seeded synthetic inputs shaped like the reference's datasets.

Real SemanticKITTI / nuScenes files are not available offline, so every test and
benchmark runs on ray-cast scans of a small static world:

* ``raycast_scan``      -- one LiDAR sweep (SK: 64 beams x 2083 azimuths, NU: 32 x 1090)
                           in the sensor frame, ``float32 [Np, 4] = (x, y, z, intensity)``.
* ``score_transform``   -- the exact voxelisation the reference applies before the network
                           (dataset/sk_dataset.py:143-169): random affine + x-flip + yaw,
                           x20 (0.05 m voxels), random shift into [0, 8192)^3, ``astype(int)``,
                           ``np.unique(axis=0)`` keeping the first point's features.
* ``collate_views``     -- batch layout of dataset/sk_dataset.py:188-242:
                           coords ``int32 [N, 4] = (x, y, z, batch)``, feats ``float32 [N, 4]``,
                           ``inverse_indices`` offset per view.
* ``make_sequence``     -- a frame sequence with poses, registered float64 coordinates
                           (dataset/prepare_kdtree_sk.py:77-80) and 20 balanced regions per
                           frame in the ``(sv_id, sv2point)`` format of
                           dataset/prepare_supervoxel_kmeans_sk.py:60-74.

Nothing here is performance critical; it is host-side numpy.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass, field

import numpy as np

SENSORS = {
    # name: (beams, elev_top_deg, elev_bottom_deg, azimuth_steps, height_m, r_min, r_max, classes)
    "SK": (64, 2.0, -24.8, 2083, 1.73, 2.0, 80.0, 19),
    "NU": (32, 10.0, -30.0, 1090, 1.84, 2.0, 70.0, 16),
}

WALL_HEIGHT = 6.0


def _wall_offset(x: np.ndarray, side: float) -> np.ndarray:
    """Static street canyon: the two building lines wobble between 8 and 16 m."""
    return side * (12.0 + 3.0 * np.sin(x / 15.0 + (0.7 if side > 0 else 2.1))
                   + 1.0 * np.sin(x / 4.0 + 1.3 * side))


def raycast_scan(seed: int, kind: str = "SK", pose: np.ndarray | None = None) -> np.ndarray:
    """One sweep in the SENSOR frame.  ``pose`` is the 4x4 float64 sensor->world matrix."""
    beams, top, bot, steps, height, rmin, rmax, _ = SENSORS[kind]
    rng = np.random.default_rng(seed)
    if pose is None:
        pose = np.eye(4)
        pose[2, 3] = height
    elev = np.deg2rad(np.linspace(top, bot, beams))
    azim = np.linspace(-math.pi, math.pi, steps, endpoint=False) + rng.uniform(0, 2 * math.pi / steps)
    el, az = np.meshgrid(elev, azim, indexing="ij")
    d_s = np.stack([np.cos(el) * np.cos(az), np.cos(el) * np.sin(az), np.sin(el)], -1).reshape(-1, 3)
    rot, org = pose[:3, :3], pose[:3, 3]
    d_w = d_s @ rot.T
    t_best = np.full(d_w.shape[0], np.inf)
    # ground plane z = 0
    with np.errstate(divide="ignore", invalid="ignore"):
        t_g = np.where(d_w[:, 2] < -1e-6, -org[2] / d_w[:, 2], np.inf)
    t_best = np.minimum(t_best, t_g)
    # wobbling walls y = w(x): fixed-point iteration from the planar guess
    for side in (1.0, -1.0):
        with np.errstate(divide="ignore", invalid="ignore"):
            ok = d_w[:, 1] * side > 1e-3
            t = np.where(ok, (_wall_offset(np.full_like(t_best, org[0]), side) - org[1]) / d_w[:, 1], np.inf)
            for _ in range(6):
                xw = org[0] + np.where(np.isfinite(t), t, 0.0) * d_w[:, 0]
                t = np.where(ok, (_wall_offset(xw, side) - org[1]) / d_w[:, 1], np.inf)
        zw = org[2] + np.where(np.isfinite(t), t, 0.0) * d_w[:, 2]
        t = np.where((zw >= 0.0) & (zw <= WALL_HEIGHT) & (t > 0), t, np.inf)
        t_best = np.minimum(t_best, t)
    r = t_best + rng.normal(0.0, 0.02, t_best.shape)
    keep = np.isfinite(t_best) & (r > rmin) & (r < rmax)
    pts = d_s[keep] * r[keep, None]
    inten = rng.uniform(0.0, 1.0, pts.shape[0])
    return np.concatenate([pts, inten[:, None]], 1).astype(np.float32)


def score_transform(raw: np.ndarray, rs: np.random.RandomState, scale: float = 20.0,
                    full_scale: float = 8192.0):
    """dataset/sk_dataset.py:101-104,143-169 with an explicit RandomState.

    Returns ``coords_v int64 [Nv,3]`` (lexicographically sorted), ``feats_v float32 [Nv,4]``
    (first point of each voxel) and ``inverse_idxs int64 [Np]``.
    """
    feats_p = np.zeros_like(raw)
    coords_p = raw[:, :3]
    feats_p[:, 3] = raw[:, 3]
    trans_m = np.eye(3) + rs.randn(3, 3) * 0.1
    trans_m[0][0] *= rs.randint(0, 2) * 2 - 1
    theta = rs.rand() * 2 * math.pi
    trans_m = np.matmul(trans_m, [[math.cos(theta), math.sin(theta), 0],
                                  [-math.sin(theta), math.cos(theta), 0], [0, 0, 1]])
    coords_p = np.matmul(coords_p, trans_m)
    feats_p[:, :3] = coords_p
    coords_p *= scale
    fs = np.array([full_scale] * 3)
    cmin, cmax = coords_p.min(0), coords_p.max(0)
    offset = (-cmin + np.clip(fs - cmax + cmin - 0.001, 0, None) * rs.rand(3)
              + np.clip(fs - cmax + cmin + 0.001, None, 0) * rs.rand(3))
    coords_p += offset
    assert (coords_p.min(1) >= 0).all() and (coords_p.max(1) < full_scale).all()
    coords_v = coords_p.astype(int)
    _, uniq, inv = np.unique(coords_v, axis=0, return_index=True, return_inverse=True)
    return coords_v[uniq], feats_p[uniq], inv.reshape(-1).astype(np.int64)


def collate_views(views):
    """dataset/sk_dataset.py:188-242 in 'score' mode."""
    coords, feats, inverse, off = [], [], [], 0
    for b, (c, f, inv) in enumerate(views):
        cb = np.concatenate([c.astype(np.int32), np.full((c.shape[0], 1), b, np.int32)], 1)
        coords.append(cb)
        feats.append(f.astype(np.float32))
        inverse.append(inv + off)
        off += c.shape[0]          # == max(previous inverse) + 1
    return (np.ascontiguousarray(np.concatenate(coords, 0)),
            np.ascontiguousarray(np.concatenate(feats, 0)),
            np.concatenate(inverse, 0).astype(np.int64))


def tta_batch(raw: np.ndarray, seed: int, inf_reps: int = 8):
    """The batch score/prob_inference.py:91-97 sees: ``inf_reps`` augmented views of ONE scan."""
    rs = np.random.RandomState(seed)
    return collate_views([score_transform(raw, rs) for _ in range(inf_reps)])


def scan_batch(seed: int, kind: str = "SK", batch: int = 8):
    """``batch`` different scans, one view each (BASELINE config 2 shape)."""
    rs = np.random.RandomState(seed)
    return collate_views([score_transform(raycast_scan(seed * 1000 + b, kind), rs) for b in range(batch)])


@dataclass
class Sequence:
    kind: str
    raw: list = field(default_factory=list)          # per frame float32 [Np,4] sensor frame
    poses: list = field(default_factory=list)        # per frame float64 [4,4]
    xyz: list = field(default_factory=list)          # per frame float64 [Np,3] registered
    sv_id: list = field(default_factory=list)        # per frame int64 [20]
    sv2point: list = field(default_factory=list)     # per frame list of int64 index arrays
    region_of_point: list = field(default_factory=list)  # per frame int32 [Np] (local region 0..19)

    @property
    def n_frames(self):
        return len(self.raw)


def make_pose(i: int, height: float, step: float = 1.0) -> np.ndarray:
    yaw = 0.05 * math.sin(i / 40.0)
    pose = np.eye(4)
    pose[:3, :3] = [[math.cos(yaw), -math.sin(yaw), 0], [math.sin(yaw), math.cos(yaw), 0], [0, 0, 1]]
    pose[:3, 3] = [i * step, 0.3 * math.sin(i / 25.0), height]
    return pose


def balanced_regions(raw: np.ndarray, n_regions: int = 20):
    """Stand-in for KMeansConstrained(20, +-5 %) (dependency absent): equal-count azimuth sectors."""
    order = np.argsort(np.arctan2(raw[:, 1], raw[:, 0]), kind="stable")
    label = np.empty(raw.shape[0], np.int32)
    bounds = np.linspace(0, raw.shape[0], n_regions + 1).astype(np.int64)
    for r in range(n_regions):
        label[order[bounds[r]:bounds[r + 1]]] = r
    sv2point = [np.where(label == r)[0] for r in np.unique(label)]
    return label, sv2point


def make_sequence(n_frames: int, kind: str = "SK", seed: int = 0, sv_id_start: int = 0,
                  max_points: int | None = None, step: float = 1.0) -> Sequence:
    height = SENSORS[kind][4]
    seq = Sequence(kind)
    nxt = sv_id_start
    for i in range(n_frames):
        pose = make_pose(i, height, step)
        raw = raycast_scan(seed * 100003 + i, kind, pose)
        if max_points is not None and raw.shape[0] > max_points:
            sel = np.sort(np.random.default_rng(seed + i).choice(raw.shape[0], max_points, replace=False))
            raw = raw[sel]
        coords = raw[:, :3]
        h = np.hstack((coords, np.ones_like(coords[:, :1])))
        h = np.sum(np.expand_dims(h, 2) * pose.T, axis=1)      # prepare_kdtree_sk.py:78-80
        label, sv2point = balanced_regions(raw)
        seq.raw.append(raw)
        seq.poses.append(pose)
        seq.xyz.append(np.ascontiguousarray(h[:, :3]))
        seq.sv_id.append(np.arange(len(sv2point)) + nxt)
        seq.sv2point.append(sv2point)
        seq.region_of_point.append(label)
        nxt += len(sv2point)
    return seq


def synthetic_probs(xyz: np.ndarray, n_cls: int, seed: int, noise: float = 0.6) -> np.ndarray:
    """Smooth-in-space softmax field + per-frame noise: float32 [Np, n_cls] rows summing to 1."""
    rng = np.random.default_rng(seed)
    basis = np.random.default_rng(1234).normal(size=(3, n_cls))
    logits = np.sin(xyz @ basis * 0.35) * 2.0 + rng.normal(0, noise, (xyz.shape[0], n_cls))
    logits -= logits.max(1, keepdims=True)
    e = np.exp(logits)
    return (e / e.sum(1, keepdims=True)).astype(np.float32)


def region_table(n_frames: int, seed: int = 0, regions_per_frame: int = 20, labelled_frac: float = 0.01):
    """Per-region arrays shaped like a scored driving sequence (inputs of the global selection, LiDAL.py:230-325):
    ``regions_per_frame`` azimuth sectors around an ego that advances 1 m per frame, so regions of nearby frames overlap
    within the 5 m rule.  Returns (sv_flags int, sv_interds f32, sv_interes f32, sv_pnums int, sv_centers f32 [.,3]);
    ``labelled_frac`` of the frames start fully labelled (dataset/sk_dataloader.py:99-118)."""
    rng = np.random.default_rng(seed)
    n = n_frames * regions_per_frame
    frame = np.repeat(np.arange(n_frames), regions_per_frame)
    sector = np.tile(np.arange(regions_per_frame), n_frames)
    ang = (sector + 0.5) / regions_per_frame * 2 * math.pi
    rad = rng.uniform(4.0, 18.0, n)
    centers = np.stack([frame * 1.0 + rad * np.cos(ang), rad * np.sin(ang), rng.uniform(-0.5, 1.5, n)], 1).astype(np.float32)
    interds = rng.gamma(2.0, 0.05, n).astype(np.float32)
    interds[rng.random(n) < 0.01] = 0.0                      # regions without a single matched point
    interes = rng.uniform(0.1, 2.5, n).astype(np.float32)
    pnums = rng.integers(6200, 6900, n)
    flags = np.zeros(n, int)
    lab = rng.choice(n_frames, max(1, int(round(labelled_frac * n_frames))), replace=False)
    flags[np.isin(frame, lab)] = 1
    return flags, interds, interes, pnums.astype(int), centers


def write_scoring_tree(root: str, model_name: str = "SPVCNN", lengths=None, max_points: int = 300, n_cls: int = 19, seed: int = 40,
                       dataset_name: str = "SK"):
    """A synthetic ``Processing_files/<dataset>/`` tree in the reference's formats for round r_id = 1 -- what
    ``python -m score.sv_level.LiDAL`` reads (score/sv_level/LiDAL.py:152-197): per frame a flag file
    (sv_flag/KMeans/0r/<seq>/<frame>.npy), a prob map (prob_map/<model>/fr/0r/...), a pickled sklearn KDTree of the registered
    points (kdtree/<seq>/<frame>.pickle, dataset/prepare_kdtree_sk.py:83-88) and a pickled ``(sv_id, sv2point)``
    (super_voxel/KMeans/<seq>/<frame>.pickle, dataset/prepare_supervoxel_kmeans_sk.py:60-74).  ``lengths`` maps sequence
    names to frame counts; region ids run through the dataset in that order.  Returns the number of regions."""
    import pickle
    from sklearn.neighbors import KDTree
    lengths = lengths or {"00": 26, "01": 25}
    base = os.path.join(root, "Processing_files", dataset_name)
    rng = np.random.default_rng(seed)
    nxt = 0
    for s, (seq_id, n) in enumerate(lengths.items()):
        seq = make_sequence(n, "NU", seed=seed + s, sv_id_start=nxt, max_points=max_points, step=0.6)
        nxt = int(seq.sv_id[-1][-1]) + 1
        dirs = {k: os.path.join(base, *v, seq_id) for k, v in dict(
            flag=("sv_flag", "KMeans", "0r"), prob=("prob_map", model_name, "fr", "0r"), kd=("kdtree",), sv=("super_voxel", "KMeans")).items()}
        for d in dirs.values():
            os.makedirs(d, exist_ok=True)
        for i in range(n):
            name = f"{i:06d}"
            flags = np.zeros(len(seq.sv_id[i]))
            if i == 3 * s:
                flags[:] = 1                                     # one fully labelled frame per sequence
            flags[rng.random(len(flags)) < 0.05] = 2             # stale pseudo labels (reset at LiDAL.py:286)
            np.save(os.path.join(dirs["flag"], name + ".npy"), flags)
            np.save(os.path.join(dirs["prob"], name + ".npy"), synthetic_probs(seq.xyz[i], n_cls, 7000 + 100 * s + i))
            with open(os.path.join(dirs["kd"], name + ".pickle"), "wb") as f:
                pickle.dump(KDTree(seq.xyz[i]), f)
            with open(os.path.join(dirs["sv"], name + ".pickle"), "wb") as f:
                pickle.dump((seq.sv_id[i], seq.sv2point[i]), f)
    return nxt


# ------------------------------------------------------------------------------------------- device-side generator
def raycast_scan_gpu(seed: int, kind: str, pose: np.ndarray, device):
    """``raycast_scan`` with torch ops on ``device`` (same world, same sensor model; the noise stream is torch's, so the
    points differ from the numpy version in the last digits).  Used to synthesise long sequences for the benchmark without
    minutes of host ray casting.  Returns float32 [Np, 4] on device."""
    import torch
    beams, top, bot, steps, height, rmin, rmax, _ = SENSORS[kind]
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    f64 = torch.float64
    elev = torch.deg2rad(torch.linspace(top, bot, beams, dtype=f64, device=device))
    az0 = float(torch.rand(1, generator=g, device=device, dtype=f64).item()) * 2 * math.pi / steps
    azim = torch.arange(steps, dtype=f64, device=device) * (2 * math.pi / steps) - math.pi + az0
    el, az = torch.meshgrid(elev, azim, indexing="ij")
    d_s = torch.stack([torch.cos(el) * torch.cos(az), torch.cos(el) * torch.sin(az), torch.sin(el)], -1).reshape(-1, 3)
    rot = torch.as_tensor(pose[:3, :3], dtype=f64, device=device)
    org = [float(v) for v in pose[:3, 3]]
    d_w = d_s @ rot.T
    inf = torch.full((d_w.shape[0],), float("inf"), dtype=f64, device=device)
    t_best = torch.where(d_w[:, 2] < -1e-6, -org[2] / d_w[:, 2], inf)

    def wall(x, side):
        return side * (12.0 + 3.0 * torch.sin(x / 15.0 + (0.7 if side > 0 else 2.1)) + 1.0 * torch.sin(x / 4.0 + 1.3 * side))

    for side in (1.0, -1.0):
        ok = d_w[:, 1] * side > 1e-3
        x0 = torch.full_like(t_best, org[0])
        t = torch.where(ok, (wall(x0, side) - org[1]) / d_w[:, 1], inf)
        for _ in range(6):
            xw = org[0] + torch.where(torch.isfinite(t), t, torch.zeros_like(t)) * d_w[:, 0]
            t = torch.where(ok, (wall(xw, side) - org[1]) / d_w[:, 1], inf)
        zw = org[2] + torch.where(torch.isfinite(t), t, torch.zeros_like(t)) * d_w[:, 2]
        t = torch.where((zw >= 0.0) & (zw <= WALL_HEIGHT) & (t > 0), t, inf)
        t_best = torch.minimum(t_best, t)
    r = t_best + torch.randn(t_best.shape, generator=g, device=device, dtype=f64) * 0.02
    keep = torch.isfinite(t_best) & (r > rmin) & (r < rmax)
    pts = d_s[keep] * r[keep, None]
    inten = torch.rand(pts.shape[0], generator=g, device=device, dtype=f64)
    return torch.cat([pts, inten[:, None]], 1).float().contiguous()


def balanced_regions_gpu(raw, n_regions: int = 20):
    """``balanced_regions`` on device: equal-count azimuth sectors as a CSR pair (region_ptr int32 [R+1], region_pts int32 [Np],
    point ids ascending inside a region -- the order ``np.where(label == r)[0]`` gives)."""
    import torch
    n = raw.shape[0]
    order = torch.sort(torch.atan2(raw[:, 1], raw[:, 0]), stable=True).indices
    bounds = torch.linspace(0, n, n_regions + 1, dtype=torch.float64).to(torch.int64)
    label = torch.empty(n, dtype=torch.int64, device=raw.device)
    sizes = (bounds[1:] - bounds[:-1]).to(raw.device)
    label[order] = torch.repeat_interleave(torch.arange(n_regions, device=raw.device), sizes)
    pts = torch.sort(label, stable=True).indices.to(torch.int32)
    return bounds.to(torch.int32).to(raw.device), pts.contiguous()


class GpuSequence:
    """A synthetic driving sequence generated on device, frame by frame: ``frame(fid)`` -> (raw f32 [Np,4] on device,
    pose float64 4x4, sv_id int64 [R], (region_ptr, region_pts) device CSR).  ``sv_id`` numbers regions globally across the
    sequence (dataset/prepare_supervoxel_kmeans_sk.py:67-69)."""

    def __init__(self, n_frames: int, kind: str = "SK", seed: int = 0, device="cuda", n_regions: int = 20, sv_id_start: int = 0):
        self.n_frames, self.kind, self.seed, self.device, self.n_regions, self.sv_id_start = n_frames, kind, seed, device, n_regions, sv_id_start

    def frame(self, fid: int):
        pose = make_pose(fid, SENSORS[self.kind][4])
        raw = raycast_scan_gpu(self.seed * 100003 + fid, self.kind, pose, self.device)
        sv_id = np.arange(self.n_regions, dtype=np.int64) + self.sv_id_start + fid * self.n_regions
        return raw, pose, sv_id, balanced_regions_gpu(raw, self.n_regions)
