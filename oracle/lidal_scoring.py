"""ORACLE (test infrastructure): numpy restatement of LiDAL's scoring chain on in-memory arrays.

* ``tta_tail``           -- score/prob_inference.py:100-113
* ``neighbour_ids``      -- score/sv_level/LiDAL.py:41-42
* ``score_frame``        -- score/sv_level/LiDAL.py:59-103 (same numpy / scipy / sklearn calls,
                            arrays instead of .npy / pickle files)
* ``select_regions``     -- score/sv_level/LiDAL.py:230-325 (prints removed)

PINNED: ``tests/golden/make_golden.py`` runs the reference's own ``worker_func`` and its
``__main__`` selection block on synthetic files in the build container and checks that this
restatement reproduces them exactly; the outputs are committed under ``tests/golden/``.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import numpy as np
from scipy.special import kl_div
from scipy.stats import entropy
from sklearn.neighbors import KDTree


def tta_tail(logits_v: np.ndarray, inverse_indices: np.ndarray, inf_reps: int):
    """logits_v f32 [Nv, C] (all views), inverse_indices int64 [inf_reps*Np] -> (prob f32 [Np,C], pred int64 [Np])."""
    import torch
    logits_p = torch.from_numpy(logits_v)[torch.from_numpy(inverse_indices)]
    prob = torch.softmax(logits_p, dim=1).numpy().reshape(inf_reps, -1, logits_v.shape[-1])
    prob_mean = np.mean(prob, axis=0)
    return prob_mean, np.argmax(prob_mean, axis=1)


def neighbour_ids(fid: int, n_frames: int, nei_num: int = 24):
    half = int(nei_num / 2)
    ids = [(fid - o - 1) if (fid - o - 1) >= 0 else (half + o + 1) for o in np.arange(half)]
    ids += [(fid + o + 1) if (fid + o + 1) <= (n_frames - 1) else (n_frames - 2 - half - o) for o in np.arange(half)]
    return [int(i) for i in ids]


def score_points(query_prob, query_points, nei_probs, nei_trees, dis_thresh=0.1):
    """LiDAL.py:59-81.  Returns (interd_points f64 [Np], intere_points f32 [Np], map_count f64 [Np])."""
    map_count = np.ones(query_prob.shape[0])
    interd_points = np.zeros(query_points.shape[0])
    sum_prob = query_prob.copy()
    epsilon = 0.00001
    for n_prob, n_tree in zip(nei_probs, nei_trees):
        dists, nearest_ids = n_tree.query(query_points, k=1, return_distance=True, dualtree=False, breadth_first=False)
        dists = dists.squeeze()
        nearest_ids = nearest_ids.squeeze()
        match_mask = dists <= dis_thresh
        sum_prob[match_mask] += n_prob[nearest_ids][match_mask]
        interd_points[match_mask] += np.sum(kl_div(query_prob[match_mask] + epsilon,
                                                   n_prob[nearest_ids][match_mask] + epsilon), axis=1)
        map_count[match_mask] += 1
    sum_prob /= np.expand_dims(map_count, 1)
    intere_points = entropy(sum_prob, axis=1)
    map_count = map_count - 1
    map_mask = map_count > 0
    interd_points[map_mask] /= map_count[map_mask]
    return interd_points, intere_points, map_count


def reduce_regions(interd_points, intere_points, query_points, sv_id, sv2point):
    """LiDAL.py:87-103 with sv_pre == False."""
    sv_interds = np.zeros_like(sv_id, dtype=np.float32)
    sv_interes = np.zeros_like(sv_id, dtype=np.float32)
    sv_pnums = np.zeros_like(sv_id, dtype=int)
    sv_centers = np.zeros((len(sv_id), 3), dtype=np.float32)
    for sv_idx, p_ids in enumerate(sv2point):
        sv_pnums[sv_idx] = len(p_ids)
        sv_centers[sv_idx] = query_points[p_ids].mean(0)
        sv_interds[sv_idx] = interd_points[p_ids].mean()
        sv_interes[sv_idx] = intere_points[p_ids].mean()
    return sv_id, sv_interds, sv_interes, sv_pnums, sv_centers


def score_frame(fid, probs, xyz, trees, sv_id, sv2point, nei_num=24, dis_thresh=0.1):
    """One ``worker_func(id)`` call; ``probs``/``xyz``/``trees`` are per-frame lists for the sequence."""
    nids = neighbour_ids(fid, len(probs), nei_num)
    d, e, _ = score_points(probs[fid], np.asarray(trees[fid].data), [probs[n] for n in nids],
                           [trees[n] for n in nids], dis_thresh)
    return reduce_regions(d, e, np.asarray(trees[fid].data), sv_id, sv2point)


def build_trees(xyz_list):
    """dataset/prepare_kdtree_sk.py:83."""
    return [KDTree(x) for x in xyz_list]


def select_regions(sv_flags, sv_interds, sv_interes, sv_pnums, sv_centers, train_point_num, sv_dis_thresh=5.0):
    """LiDAL.py:230-325.  ``sv_flags`` is the float64 array np.append builds; returns int flags (0/1/2)."""
    sv_flags = sv_flags.astype(int)
    unlabeled_ids = np.where(sv_flags == 0)[0]
    unlabeled_interds = sv_interds[unlabeled_ids]
    sorted_ids = np.argsort(unlabeled_interds)
    added_ids = set()
    point_limit = round(0.01 * train_point_num)
    for idx in reversed(sorted_ids):
        sv_id = unlabeled_ids[idx]
        sv_c = sv_centers[sv_id]
        flag = True
        for l_sv_id in added_ids:
            l_sv_c = sv_centers[l_sv_id]
            dist = np.sqrt(np.square(sv_c - l_sv_c).sum())
            if dist < sv_dis_thresh:
                flag = False
                if sv_interes[l_sv_id] < sv_interes[sv_id]:
                    sv_flags[sv_id] = 1
                    sv_flags[l_sv_id] = 0
                    added_ids.add(sv_id)
                    added_ids.remove(l_sv_id)
                    point_limit = point_limit + sv_pnums[l_sv_id] - sv_pnums[sv_id]
                break
        if flag:
            point_limit -= sv_pnums[sv_id]
            if point_limit < 0:
                break
            sv_flags[sv_id] = 1
            added_ids.add(sv_id)

    unlabeled_ids = np.where(sv_flags == 0)[0]
    unlabeled_interds = sv_interds[unlabeled_ids]
    sorted_ids = np.argsort(unlabeled_interds)
    sv_flags[sv_flags == 2] = 0
    added_ids = set()
    point_limit = round(0.01 * train_point_num)
    for idx in sorted_ids:
        if unlabeled_interds[idx] == 0:
            continue
        sv_id = unlabeled_ids[idx]
        sv_c = sv_centers[sv_id]
        flag = True
        for l_sv_id in added_ids:
            l_sv_c = sv_centers[l_sv_id]
            dist = np.sqrt(np.square(sv_c - l_sv_c).sum())
            if dist < sv_dis_thresh:
                flag = False
                if sv_interes[l_sv_id] > sv_interes[sv_id]:
                    sv_flags[sv_id] = 2
                    sv_flags[l_sv_id] = 0
                    added_ids.add(sv_id)
                    added_ids.remove(l_sv_id)
                    point_limit = point_limit + sv_pnums[l_sv_id] - sv_pnums[sv_id]
                break
        if flag:
            point_limit -= sv_pnums[sv_id]
            if point_limit < 0:
                break
            sv_flags[sv_id] = 2
            added_ids.add(sv_id)
    return sv_flags


def score_dataset(sequences, nei_num=24, dis_thresh=0.1, n_regions_total=None, sv_pnums=None, sv_centers=None):
    """LiDAL.py:163-222 on the reference's file formats: per sequence (in train_split order) every frame's ``worker_func``
    result is scattered into the global arrays by ``sv_id``; centres get ``idx * 1000.0`` (:218); cached ``sv_pnums`` /
    ``sv_centers`` are reused when given (``sv_pre``, :171-175).  ``sequences``: [(prob_files, kdtree_files, sv_info_files)].
    Returns (sv_interds f32, sv_interes f32, sv_pnums int, sv_centers f32 [.,3], sv_pre)."""
    import pickle
    sv_pre = sv_pnums is not None and sv_centers is not None
    sv_interds = np.zeros(n_regions_total, np.float32)
    sv_interes = np.zeros(n_regions_total, np.float32)
    if not sv_pre:
        sv_pnums = np.zeros(n_regions_total, int)
        sv_centers = np.zeros((n_regions_total, 3), np.float32)
    for idx, (prob_files, kdtree_files, sv_info_files) in enumerate(sequences):
        assert len(prob_files) == len(kdtree_files) == len(sv_info_files)
        probs = [np.load(p) for p in prob_files]
        trees = []
        for k in kdtree_files:
            with open(k, "rb") as f:
                trees.append(pickle.load(f))
        xyz = [np.asarray(t.data) for t in trees]
        for fid, s in enumerate(sv_info_files):
            with open(s, "rb") as f:
                sv_id, sv2point = pickle.load(f)
            ids, d, e, pn, c = score_frame(fid, probs, xyz, trees, sv_id, sv2point, nei_num, dis_thresh)
            sv_interds[ids], sv_interes[ids] = d, e
            if not sv_pre:
                sv_pnums[ids] = pn
                sv_centers[ids] = c + idx * 1000.0
    return sv_interds, sv_interes, sv_pnums, sv_centers, sv_pre
