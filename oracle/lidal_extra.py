"""CPU restatement (TEST INFRASTRUCTURE ONLY -- see oracle/README.md) of the reference steps either side of the network
that lidal_b200 also runs on device.  Each function follows the cited reference lines operation by operation; all of them
are pinned against the reference's own code run in the build container (tests/golden/make_golden.py -> extra.npz,
voxelizer.npz).

  register_points     dataset/prepare_kdtree_sk.py:76-80
  score_batch         dataset/sk_dataset.py:101-104,143-169 (``__getitem__`` in 'score' mode) + collate_fn :188-242
  outfeat_mean        score/prob_inference.py:103-105,116-118
  segment_entropy     score/frame_level/segment_entropy.py:41-48
  redal_worker        score/sv_level/ReDAL.py:60-84
"""
from __future__ import annotations

import math

import numpy as np


def register_points(raw: np.ndarray, pose: np.ndarray) -> np.ndarray:
    """dataset/prepare_kdtree_sk.py:72-80: raw float32 [N,4], pose float64 [4,4] -> registered float64 [N,3]."""
    coords = raw[:, :3]
    hcoords = np.hstack((coords, np.ones_like(coords[:, :1])))
    hcoords = np.sum(np.expand_dims(hcoords, 2) * pose.T, axis=1)
    return hcoords[:, :3]


def score_view(raw: np.ndarray, rs, scale=20, full_scale=(8192, 8192, 8192)):
    """dataset/sk_dataset.py:101-104,143-169 for one item in 'score' mode; ``rs`` stands for the global ``np.random``."""
    raw_data = raw
    feats_p = np.zeros_like(raw_data)
    coords_p = raw_data[:, :3]
    feats_p[:, 3] = raw_data[:, 3]
    trans_m = np.eye(3) + rs.randn(3, 3) * 0.1
    trans_m[0][0] *= rs.randint(0, 2) * 2 - 1
    theta = rs.rand() * 2 * math.pi
    trans_m = np.matmul(trans_m, [[math.cos(theta), math.sin(theta), 0], [-math.sin(theta), math.cos(theta), 0], [0, 0, 1]])
    coords_p = np.matmul(coords_p, trans_m)
    feats_p[:, :3] = coords_p
    coords_p *= scale
    coords_min = coords_p.min(0)
    coords_max = coords_p.max(0)
    offset = -coords_min + np.clip(full_scale - coords_max + coords_min - 0.001, 0, None) * rs.rand(3) \
        + np.clip(full_scale - coords_max + coords_min + 0.001, None, 0) * rs.rand(3)
    coords_p += offset
    valid_idxs = (coords_p.min(1) >= 0) * (coords_p.max(1) < full_scale[0])
    assert sum(valid_idxs) == len(valid_idxs), 'input voxels are not valid'
    coords_v = coords_p.astype(int)
    _, unique_idxs, inverse_idxs = np.unique(coords_v, axis=0, return_index=True, return_inverse=True)
    return coords_v[unique_idxs], feats_p[unique_idxs], np.asarray(inverse_idxs).reshape(-1)


def score_batch(raw: np.ndarray, seed: int, inf_reps: int = 8):
    """The batch score/prob_inference.py:91-97 receives: ``inf_reps`` items of the same scan through collate_fn (:188-242).
    Returns coords int32 [N,4] (x,y,z,batch), feats float32 [N,4], inverse_indices int64 [inf_reps*Np]."""
    rs = np.random.RandomState(seed)
    coords, feats, inverse, off = [], [], [], 0
    for b in range(inf_reps):
        c, f, inv = score_view(raw, rs)
        coords.append(np.concatenate([c.astype(np.int32), np.full((c.shape[0], 1), b, np.int32)], 1))
        feats.append(f.astype(np.float32))
        inverse.append(inv.astype(np.int64) + off)
        off = int(inverse[-1].max()) + 1                      # collate_fn: max(previous) + 1
    return np.concatenate(coords, 0), np.concatenate(feats, 0), np.concatenate(inverse, 0)


def outfeat_mean(out_feat_v: np.ndarray, inverse_indices: np.ndarray, inf_reps: int) -> np.ndarray:
    """score/prob_inference.py:103-105,116-118: gather by inverse index, reshape to views, float32 mean over views."""
    out_feat_p = out_feat_v[inverse_indices]
    out_feat = out_feat_p.reshape(inf_reps, -1, out_feat_p.shape[-1])
    return np.mean(out_feat, axis=0)


def segment_entropy(pred: np.ndarray, sv2point, class_num: int) -> float:
    """score/frame_level/segment_entropy.py:41-48."""
    frame_sege = 0.0
    for sv_idx, p_ids in enumerate(sv2point):
        sv_preds = pred[p_ids]
        sv_sege = 0.0
        for class_id in range(class_num):
            q_c = (sv_preds == class_id).sum() / sv_preds.shape[0]
            sv_sege += -q_c * np.log2(q_c + 1e-12)
        frame_sege += sv_sege * sv_preds.shape[0] / pred.shape[0]
    return frame_sege


REDAL_ALPHA, REDAL_GAMMA, REDAL_FT_DIM = 1.0, 0.05, 96        # score/sv_level/ReDAL.py:15,17,21


def redal_worker(prob, outfeat, curvature, sv_id, sv2point, sv_pre: bool):
    """score/sv_level/ReDAL.py:60-84 on in-memory arrays (curvature already float32)."""
    uncertain = np.mean(-prob * np.log2(prob + 1e-12), axis=1)
    point_score = REDAL_ALPHA * uncertain + REDAL_GAMMA * curvature
    sv_scores = np.zeros_like(sv_id, dtype=np.float32)
    sv_feats = np.zeros((sv_id.shape[0], REDAL_FT_DIM), dtype=np.float32)
    if not sv_pre:
        sv_pnums = np.zeros_like(sv_id, dtype=int)
    for sv_idx, p_ids in enumerate(sv2point):
        if not sv_pre:
            sv_pnums[sv_idx] = len(p_ids)
        sv_scores[sv_idx] = point_score[p_ids].mean()
        sv_feats[sv_idx] = outfeat[p_ids].mean(0)
    if not sv_pre:
        return sv_id, sv_scores, sv_feats, sv_pnums
    return sv_id, sv_scores, sv_feats
