"""Multi-GPU orchestration of the scoring chain: prob_inference -> inter-frame scoring -> selection.

One process per GPU (``torch.distributed``; NCCL on the box, gloo in the CPU tests).  The path shards by frame with no
data-path collective; the only exchanges are

* a point-to-point fetch of the neighbour-window frames a rank does not own (the +-12 frame "halo", including the
  reflected windows at the sequence ends, score/sv_level/LiDAL.py:41-42), and
* ONE all_gather of the per-region scores before the global greedy selection (LiDAL.py:208-218 -> :230).

``frame_shard`` is the reference's split (dataset/sk_dataloader.py:196-198): contiguous ceil(N / G) chunks.
The scorer is injected so the same orchestration runs on the CUDA scorer (lidal_b200.score) and, in the world-size-2
gloo test, on the CPU oracle.
"""
from __future__ import annotations

import math
from typing import Callable, Sequence

import numpy as np
import torch
import torch.distributed as dist

from .score import neighbour_ids


def frame_shard(n_frames: int, world: int, rank: int) -> range:
    """dataset/sk_dataloader.py:196-198."""
    if world <= 1:
        return range(n_frames)
    split = int(math.ceil(n_frames / world))
    return range(min(rank * split, n_frames), min((rank + 1) * split, n_frames))


def owner_of(fid: int, n_frames: int, world: int) -> int:
    return 0 if world <= 1 else fid // int(math.ceil(n_frames / world))


def needed_frames(own: range, n_frames: int, nei_num: int = 24) -> list[int]:
    """Frames whose (xyz, prob) a rank must hold to score its own frames: own + every neighbour window."""
    need = set(own)
    for f in own:
        need.update(neighbour_ids(f, n_frames, nei_num))
    return sorted(need)


def _p2p(tensor_of: Callable[[int], torch.Tensor], alloc: Callable[[int], torch.Tensor], sends, recvs, group=None):
    """sends: [(dst_rank, fid)], recvs: [(src_rank, fid)] -> {fid: tensor}.  One batched isend/irecv round."""
    ops, out = [], {}
    for dst, fid in sends:
        ops.append(dist.P2POp(dist.isend, tensor_of(fid).contiguous(), dst, group=group, tag=fid))
    for src, fid in recvs:
        out[fid] = alloc(fid)
        ops.append(dist.P2POp(dist.irecv, out[fid], src, group=group, tag=fid))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return out


def exchange_halo(local: dict, n_points: Sequence[int], n_cls: int, n_frames: int, nei_num: int = 24, device="cpu"):
    """local: {fid: (xyz f64 [Np,3], prob f32 [Np,C])} for the frames this rank owns.  Returns the same dict extended
    with every frame of its neighbour windows, fetched from the owning ranks.  Everyone can compute everyone's needs,
    so no request messages are necessary."""
    world, rank = (dist.get_world_size(), dist.get_rank()) if dist.is_initialized() else (1, 0)
    if world == 1:
        return dict(local)
    shards = [frame_shard(n_frames, world, r) for r in range(world)]
    needs = [needed_frames(shards[r], n_frames, nei_num) for r in range(world)]
    sends = [(r, f) for r in range(world) if r != rank for f in needs[r] if f in shards[rank]]
    recvs = [(owner_of(f, n_frames, world), f) for f in needs[rank] if f not in shards[rank]]
    got_xyz = _p2p(lambda f: local[f][0], lambda f: torch.empty((n_points[f], 3), dtype=torch.float64, device=device),
                   sends, recvs)
    got_prob = _p2p(lambda f: local[f][1], lambda f: torch.empty((n_points[f], n_cls), dtype=torch.float32, device=device),
                    sends, recvs)
    out = dict(local)
    for f in got_xyz:
        out[f] = (got_xyz[f], got_prob[f])
    return out


def gather_region_scores(sv_id, sv_d, sv_e, sv_n, sv_c, n_regions_total: int, device="cpu"):
    """The one collective of the path: every rank contributes the regions of its frames; all ranks obtain the global
    arrays indexed by sv_id (LiDAL.py:208-218).  Payload: 8 + 4 + 4 + 8 + 12 bytes per region."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    pack = torch.zeros((len(sv_id), 6), dtype=torch.float64, device=device)
    pack[:, 0] = torch.as_tensor(np.asarray(sv_id), dtype=torch.float64)
    pack[:, 1] = torch.as_tensor(np.asarray(sv_d), dtype=torch.float64)
    pack[:, 2] = torch.as_tensor(np.asarray(sv_e), dtype=torch.float64)
    pack[:, 3] = torch.as_tensor(np.asarray(sv_n), dtype=torch.float64)
    centres = torch.as_tensor(np.asarray(sv_c), dtype=torch.float64).reshape(-1, 3)
    if world > 1:
        counts = torch.zeros(world, dtype=torch.int64, device=device)
        counts[dist.get_rank()] = len(sv_id)
        dist.all_reduce(counts)
        cap = int(counts.max().item())
        buf = torch.zeros((cap, 9), dtype=torch.float64, device=device)
        buf[: len(sv_id), :6] = pack
        buf[: len(sv_id), 6:] = centres
        parts = [torch.empty((cap, 9), dtype=torch.float64, device=device) for _ in range(world)]
        dist.all_gather(parts, buf)
        rows = torch.cat([parts[r][: int(counts[r])] for r in range(world)]).cpu().numpy()
    else:
        rows = torch.cat([pack, torch.zeros((len(sv_id), 0), dtype=torch.float64, device=device)], 1).cpu().numpy()
        rows = np.concatenate([rows[:, :6], centres.cpu().numpy()], 1)
    ids = rows[:, 0].astype(np.int64)
    sv_interds = np.zeros(n_regions_total, np.float32)
    sv_interes = np.zeros(n_regions_total, np.float32)
    sv_pnums = np.zeros(n_regions_total, int)
    sv_centers = np.zeros((n_regions_total, 3), np.float32)
    sv_interds[ids] = rows[:, 1].astype(np.float32)          # values were float32 before packing: exact round trip
    sv_interes[ids] = rows[:, 2].astype(np.float32)
    sv_pnums[ids] = rows[:, 3].astype(int)
    sv_centers[ids] = rows[:, 6:9].astype(np.float32)
    return sv_interds, sv_interes, sv_pnums, sv_centers


def score_sequence_sharded(frames: dict, n_points, n_cls, n_frames, regions: dict, score_frame: Callable, n_regions_total,
                           nei_num=24, device="cpu"):
    """frames: {fid: (xyz, prob)} owned by this rank; regions: {fid: (sv_id, sv2point)} for owned frames.
    score_frame(fid, held_frames, sv_id, sv2point) -> (sv_id, d, e, pnums, centres) for ONE frame (the worker_func
    contract).  Returns the global per-region arrays on every rank."""
    held = exchange_halo(frames, n_points, n_cls, n_frames, nei_num, device)
    ids, ds, es, ns, cs = [], [], [], [], []
    for fid in sorted(frames):
        sv_id, d, e, pn, c = score_frame(fid, held, *regions[fid])
        ids.append(np.asarray(sv_id)); ds.append(d); es.append(e); ns.append(pn); cs.append(c)
    cat = lambda xs, shape: np.concatenate(xs) if xs else np.zeros(shape)   # noqa: E731
    return gather_region_scores(cat(ids, 0), cat(ds, 0), cat(es, 0), cat(ns, 0), cat(cs, (0, 3)), n_regions_total, device)


def cuda_score_frame(n_frames: int, nei_num=24, dis_thresh=0.1, device="cuda"):
    """score_frame callback for score_sequence_sharded on the CUDA scorer: frames (own + halo) become resident once."""
    from .score import SequenceScorer
    scorer = SequenceScorer(device, nei_num, dis_thresh, n_total=n_frames)

    def score(fid, held, sv_id, sv2point):
        for f, (xyz, prob) in held.items():
            if f not in scorer.frames:
                scorer.add_frame(xyz, prob, fid=f)
        scorer.set_regions(scorer.frames[fid], sv_id, sv2point)
        return scorer.score_frame(fid)
    return score


def infer_and_score_sequence(engine, raws, xyz, regions, seed=0, inf_reps=8, nei_num=24, dis_thresh=0.1, device="cuda"):
    """The whole LiDAL chain for one sequence on ONE GPU, nothing leaves the device between stages:
    raw scan -> 8 TTA views (voxelizer, F1) -> network (engine) -> softmax / mean / argmax (tta_tail) -> resident prob map
    -> inter-frame divergence / entropy against the +-12 frame window -> per-region means.
    raws: per-frame float32 [Np,4] sensor-frame points (host or device); xyz: per-frame float64 [Np,3] registered coordinates;
    regions: per-frame (sv_id, sv2point).  Returns (per-frame worker_func tuples, timings dict in ms)."""
    from . import score, voxelizer
    dev = torch.device(device)
    ev = lambda: torch.cuda.Event(enable_timing=True)          # noqa: E731
    t0, t1, t2 = ev(), ev(), ev()
    scorer = score.SequenceScorer(dev, nei_num, dis_thresh)
    t0.record()
    for i, raw in enumerate(raws):
        raw_dev = torch.as_tensor(raw).to(dev)
        coords, feats, inverse = voxelizer.tta_batch_gpu(raw_dev, seed=seed + i, inf_reps=inf_reps)
        logits = engine(coords, feats)
        prob, _pred = score.tta_tail(logits, inverse, inf_reps)
        scorer.add_frame(xyz[i], prob, *regions[i])
    t1.record()
    out = [scorer.score_frame(i) for i in range(len(raws))]
    t2.record()
    torch.cuda.synchronize()
    n = len(raws)
    return out, {"prob_inference_ms_per_frame": t0.elapsed_time(t1) / n, "scoring_ms_per_frame": t1.elapsed_time(t2) / n,
                 "frames_per_sec": 1e3 * n / t0.elapsed_time(t2), "frames": n}
