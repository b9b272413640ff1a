"""Multi-GPU orchestration of the scoring chain: prob_inference -> inter-frame scoring -> selection.

One process per GPU (``torch.distributed``; NCCL on the box, gloo in the CPU tests).  The path shards by frame with no
data-path collective; the only exchanges are

* a point-to-point fetch of the neighbour-window frames a rank does not own (the +-12 frame "halo", including the
  reflected windows at the sequence ends, score/sv_level/LiDAL.py:41-42), and
* ONE all_gather of the per-region scores before the global greedy selection (LiDAL.py:208-218 -> :230).

``frame_shard`` is the reference's split (dataset/sk_dataloader.py:196-198): contiguous ceil(N / G) chunks.
A dataset of many short sequences (nuScenes: 850 scenes of ~40 frames, BASELINE configs[4]) is sharded by WHOLE sequences
instead (``pack_sequences``: sequences are independent, LiDAL.py:185, so there is no halo at all), and the same single
all_gather follows.
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


def pack_sequences(frame_counts: Sequence[int], world: int) -> list[list[int]]:
    """Whole sequences -> ranks (many-sequence datasets): longest-processing-time greedy bin packing by
    frame count -- sequences in descending length (ties: lower index first) each go to the currently least loaded rank (ties:
    lower rank).  Deterministic, so every rank computes the same assignment without communication.  Returns, per rank, its
    sequence indices in ascending order (= ``train_split`` order, the order LiDAL.py:185 visits them in)."""
    world = max(int(world), 1)
    load = [0] * world
    bins: list[list[int]] = [[] for _ in range(world)]
    for idx in sorted(range(len(frame_counts)), key=lambda i: (-int(frame_counts[i]), i)):
        r = min(range(world), key=lambda q: (load[q], q))
        bins[r].append(idx)
        load[r] += int(frame_counts[idx])
    return [sorted(b) for b in bins]


def sequence_region_offsets(region_counts: Sequence[int]) -> list[int]:
    """First global region id of every sequence when ``sv_id`` runs through the dataset in ``train_split`` order
    (dataset/prepare_supervoxel_kmeans_sk.py:67-69 numbers regions across frames and sequences without gaps)."""
    out, acc = [], 0
    for n in region_counts:
        out.append(acc)
        acc += int(n)
    return out


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


def gather_region_scores(sv_id, sv_d, sv_e, sv_n, sv_c, n_regions_total: int, device="cpu", timings: dict | None = None):
    """The one collective of the path: every rank contributes the regions of its frames; all ranks obtain the global
    arrays indexed by sv_id (LiDAL.py:208-218).  Payload: 9 float64 per region (id, d, e, pnum, centre, padded to a common
    row count); inputs may be numpy arrays or tensors already on ``device`` (then nothing touches the host before the
    all_gather).  float32 / int values below 2^53 survive the float64 packing exactly."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    as_t = lambda x: (x if torch.is_tensor(x) else torch.as_tensor(np.asarray(x))).to(device=device, dtype=torch.float64)  # noqa: E731
    n_loc = len(sv_id)
    pack = torch.zeros((n_loc, 9), dtype=torch.float64, device=device)
    if n_loc:
        pack[:, 0], pack[:, 1], pack[:, 2], pack[:, 3] = as_t(sv_id), as_t(sv_d), as_t(sv_e), as_t(sv_n)
        pack[:, 6:] = as_t(sv_c).reshape(-1, 3)
    if world > 1:
        ev0 = ev1 = None
        if timings is not None and pack.is_cuda:
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ev0.record()
        counts = torch.zeros(world, dtype=torch.int64, device=device)
        counts[dist.get_rank()] = n_loc
        dist.all_reduce(counts)
        counts = counts.cpu()
        cap = int(counts.max().item())
        buf = torch.zeros((cap, 9), dtype=torch.float64, device=device)
        buf[:n_loc] = pack
        parts = [torch.empty((cap, 9), dtype=torch.float64, device=device) for _ in range(world)]
        dist.all_gather(parts, buf)
        if ev0 is not None:
            ev1.record()
            torch.cuda.synchronize()
            timings["all_gather_ms"] = ev0.elapsed_time(ev1)
            timings["all_gather_bytes_per_rank"] = cap * 9 * 8
        rows = torch.cat([parts[r][: int(counts[r])] for r in range(world)]).cpu().numpy()
    else:
        rows = pack.cpu().numpy()
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


def score_sequences_sharded(frame_counts: Sequence[int], score_sequence: Callable, n_regions_total: int, device="cpu",
                            timings: dict | None = None):
    """Many-sequence datasets (LiDAL.py:185-218 over ``train_split``): whole sequences are packed onto the ranks
    (``pack_sequences``), each rank scores its sequences with no exchange at all, adds the reference's ``idx * 1000.0``
    to the region centres of sequence ``idx`` (LiDAL.py:218; float32 + Python float stays float32) and the per-region
    rows meet in the one all_gather.  ``score_sequence(idx) -> (sv_id, d, e, pnums, centres)`` for ALL frames of sequence
    ``idx`` (numpy arrays or tensors on ``device``).  Returns the global per-region arrays on every rank."""
    world, rank = (dist.get_world_size(), dist.get_rank()) if dist.is_initialized() else (1, 0)
    mine = pack_sequences(frame_counts, world)[rank]
    ids, ds, es, ns, cs = [], [], [], [], []
    for idx in mine:
        sv_id, d, e, pn, c = score_sequence(idx)
        if torch.is_tensor(c):
            c = c.to(torch.float32) + idx * 1000.0
        else:
            c = np.asarray(c, np.float32) + idx * 1000.0
        ids.append(sv_id); ds.append(d); es.append(e); ns.append(pn); cs.append(c)
    if ids and all(torch.is_tensor(x) for x in ids + ds + es + ns + cs):
        packed = [torch.cat(xs) for xs in (ids, ds, es, ns, cs)]
    else:
        as_np = lambda x: x.cpu().numpy() if torch.is_tensor(x) else np.asarray(x)      # noqa: E731
        cat = lambda xs, shape: np.concatenate([as_np(x) for x in xs]) if xs else np.zeros(shape)   # noqa: E731
        packed = [cat(ids, 0), cat(ds, 0), cat(es, 0), cat(ns, 0), cat(cs, (0, 3))]
    return gather_region_scores(*packed, n_regions_total, device, timings)


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


def prob_inference_frame(engine, raw_dev, seed, inf_reps=8, want_feat=False):
    """score/prob_inference.py:91-118 for ONE frame, everything on device: the ``inf_reps`` TTA views (voxelizer, F1) ->
    network (engine) -> softmax / mean / argmax (tta_tail).  Returns (prob f32 [Np,C], pred i64 [Np][, out_feat f32 [Np,96]])."""
    from . import score, voxelizer
    coords, feats, inverse = voxelizer.tta_batch_gpu(raw_dev, seed=seed, inf_reps=inf_reps)
    if want_feat:
        logits, feat = engine(coords, feats, return_feat=True)
        return score.tta_tail(logits, inverse, inf_reps, out_feat=feat)
    return score.tta_tail(engine(coords, feats), inverse, inf_reps)


def run_sequence_sharded(engine, frame_source: Callable, n_frames: int, n_cls: int, n_regions_total: int, seed=0, inf_reps=8,
                         nei_num=24, dis_thresh=0.1, device="cuda", select_with=None, keep_scorer=False):
    """BASELINE config 3 / 5 for ONE sequence on the ranks of the default process group (or one GPU):

      frames shard as dataset/sk_dataloader.py:196-198 -> per own frame: H2D of the raw scan, pose registration (F2), 8-view
      prob_inference, resident prob map + hash grid  ->  halo: the +-12 frame windows a rank does not own arrive by P2P
      (xyz f64 + prob f32 per frame)  ->  inter-frame scoring + per-region means of the own frames  ->  ONE all_gather of
      the region scores  ->  (optional) the global selection, replicated on every rank (it is sequential and tiny).

    frame_source(fid) -> (raw f32 [Np,4] host (pinned) or device, pose 4x4 float64, sv_id int64 [R], regions) with regions the
    reference's ``sv2point`` lists or a device CSR pair.  Pinned host scans are uploaded on the side stream and overlap the
    previous frame's network; a device-resident scan makes the side stream wait for the caller's stream first (no overlap).  select_with = (sv_flags, train_point_num) runs ``select_regions``.
    Returns (sv_interds, sv_interes, sv_pnums, sv_centers, flags or None, timings dict [ms, device events; host for selection])."""
    import time
    from . import score
    dev = torch.device(device)
    world, rank = (dist.get_world_size(), dist.get_rank()) if dist.is_initialized() else (1, 0)
    own = frame_shard(n_frames, world, rank)
    ev = lambda: torch.cuda.Event(enable_timing=True)          # noqa: E731
    t = [ev() for _ in range(5)]
    scorer = score.SequenceScorer(dev, nei_num, dis_thresh, n_total=n_frames)
    timings: dict = {"frames_total": n_frames, "frames_own": len(own), "world": world}
    from . import voxelizer
    from .engine import StreamPipeline
    sp = StreamPipeline(engine)
    main = torch.cuda.current_stream(dev)
    sp.prep_stream.wait_stream(main)
    t[0].record()
    for fid in own:
        raw, pose, sv_id, regions = frame_source(fid)
        if torch.is_tensor(raw) and raw.is_cuda:
            sp.prep_stream.wait_stream(main)                     # a device-resident scan was produced on the caller's stream
        # coordinate-only work of frame i (upload, registration, TTA voxelizer, kernel maps: all the host round trips) on the
        # side stream, while the GPU is still busy with the network of frame i-1 on the main stream
        with torch.cuda.stream(sp.prep_stream):
            raw_dev = torch.as_tensor(raw).to(dev, non_blocking=True)
            xyz = score.register_points(raw_dev, pose)
            coords, feats, inverse = voxelizer.tta_batch_gpu(raw_dev, seed=seed + fid, inf_reps=inf_reps)
        pr = sp.prepare(coords, feats, wait_main=False)
        main.wait_event(pr.ready)
        prob, _pred = score.tta_tail(engine.forward(pr), inverse, inf_reps)
        fr = scorer.add_frame(xyz, prob, sv_id, regions, fid=fid)
        sp.retire(pr, raw_dev, coords, feats, inverse)          # side-stream allocations stay alive until this frame's kernels are done
        if fid == own[0] and len(own) > 8:
            # The resident set grows by ~20 MB per frame (prob map, coordinates, hash grid).  Taking it from the driver one
            # cudaMalloc at a time would stall the pipeline every few frames (cudaMalloc synchronises), so the caching
            # allocator is primed once with the whole sequence's worth and hands out pieces of it afterwards.
            per_frame = xyz.numel() * 8 + prob.numel() * 4 + fr.grid.numel() + 2 * (1 << 20)
            with torch.cuda.stream(main):
                reserve = torch.empty(int(per_frame * (len(own) + 2 * nei_num) * 1.05), dtype=torch.uint8, device=dev)
                del reserve
    t[1].record()
    # ---- halo: neighbour-window frames owned by other ranks
    if world > 1:
        split = int(math.ceil(n_frames / world))
        mine = torch.zeros(split, dtype=torch.int64, device=dev)
        for j, fid in enumerate(own):
            mine[j] = scorer.frames[fid].n
        allc = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allc, mine)                              # point counts of every frame (sizes of the halo messages)
        n_points = torch.cat(allc).cpu().tolist()[:n_frames]
        local = {fid: (scorer.frames[fid].xyz, scorer.frames[fid].prob) for fid in own}
        held = exchange_halo(local, n_points, n_cls, n_frames, nei_num, dev)
        halo_bytes = 0
        for fid, (xyz, prob) in held.items():
            if fid not in scorer.frames:
                scorer.add_frame(xyz, prob, fid=fid)
                halo_bytes += xyz.numel() * 8 + prob.numel() * 4
        timings["halo_frames_received"] = len(held) - len(own)
        timings["halo_bytes_received"] = halo_bytes
    t[2].record()
    ids, d, e, pn, c = scorer.score_frames_device(list(own))
    t[3].record()
    out = gather_region_scores(ids, d, e, pn, c, n_regions_total, dev, timings)
    t[4].record()
    torch.cuda.synchronize(dev)
    timings.update(prob_inference_ms=t[0].elapsed_time(t[1]), halo_ms=t[1].elapsed_time(t[2]), scoring_ms=t[2].elapsed_time(t[3]),
                   gather_ms=t[3].elapsed_time(t[4]), device_total_ms=t[0].elapsed_time(t[4]))
    flags = None
    if select_with is not None:
        h0 = time.perf_counter()
        flags = score.select_regions(select_with[0], out[0], out[1], out[2], out[3], select_with[1], device=dev, timings=timings)
        timings["selection_ms"] = (time.perf_counter() - h0) * 1e3
    if keep_scorer:
        timings["scorer"] = scorer
    return out + (flags, timings)


def run_dataset_sharded(engine, sequence_source: Callable, frame_counts: Sequence[int], n_cls: int, n_regions_total: int,
                        seed=0, inf_reps=8, nei_num=24, dis_thresh=0.1, device="cuda", select_with=None):
    """BASELINE configs[4] (nuScenes-shaped: many short sequences) on the ranks of the default process group (or one GPU):

      whole sequences are packed onto the ranks (``pack_sequences``)  ->  per own sequence, per frame: H2D of the raw scan,
      pose registration, 8-view prob_inference, resident prob map + hash grid; then inter-frame scoring + region means of
      the sequence and its frames are dropped  ->  ``idx * 1000.0`` on the centres (LiDAL.py:218)  ->  ONE all_gather of the
      region scores  ->  (optional) the global selection, replicated.

    ``sequence_source(idx) -> frame_source`` with ``frame_source(fid)`` as for ``run_sequence_sharded``.  Sequences do not
    interact (LiDAL.py:185), so there is no halo and no collective before the all_gather.
    Returns (sv_interds, sv_interes, sv_pnums, sv_centers, flags or None, timings dict [ms])."""
    import time
    from . import score, voxelizer
    from .engine import StreamPipeline
    dev = torch.device(device)
    world, rank = (dist.get_world_size(), dist.get_rank()) if dist.is_initialized() else (1, 0)
    mine = pack_sequences(frame_counts, world)[rank]
    ev = lambda: torch.cuda.Event(enable_timing=True)          # noqa: E731
    t0, t1, t2 = ev(), ev(), ev()
    timings: dict = {"sequences_total": len(frame_counts), "sequences_own": len(mine), "world": world,
                     "frames_total": int(sum(frame_counts)), "frames_own": int(sum(frame_counts[i] for i in mine))}
    sp = StreamPipeline(engine)
    main = torch.cuda.current_stream(dev)
    sp.prep_stream.wait_stream(main)
    score_ms = []

    def score_sequence(idx):
        n_frames = int(frame_counts[idx])
        frame_source = sequence_source(idx)
        scorer = score.SequenceScorer(dev, nei_num, dis_thresh, n_total=n_frames)
        for fid in range(n_frames):
            raw, pose, sv_id, regions = frame_source(fid)
            if torch.is_tensor(raw) and raw.is_cuda:
                sp.prep_stream.wait_stream(main)                 # a device-resident scan was produced on the caller's stream
            with torch.cuda.stream(sp.prep_stream):             # coordinate-only work on the side stream (see run_sequence_sharded)
                raw_dev = torch.as_tensor(raw).to(dev, non_blocking=True)
                xyz = score.register_points(raw_dev, pose)
                coords, feats, inverse = voxelizer.tta_batch_gpu(raw_dev, seed=seed + idx * 100003 + fid, inf_reps=inf_reps)
            pr = sp.prepare(coords, feats, wait_main=False)
            main.wait_event(pr.ready)
            prob, _pred = score.tta_tail(engine.forward(pr), inverse, inf_reps)
            scorer.add_frame(xyz, prob, sv_id, regions, fid=fid)
            sp.retire(pr, raw_dev, coords, feats, inverse)
        s0, s1 = ev(), ev()
        s0.record()
        out = scorer.score_frames_device(list(range(n_frames)))
        s1.record()
        score_ms.append((s0, s1))
        # the sequence's coordinates (allocated on the side stream), prob maps and grids go back to the allocator once the
        # scoring kernels queued above have finished
        sp.retire(list(scorer.frames.values()))
        scorer.frames.clear()
        return out

    t0.record()
    local_timings: dict = {}
    mine_out = []
    for idx in mine:
        sv_id, d, e, pn, c = score_sequence(idx)
        mine_out.append((torch.as_tensor(np.asarray(sv_id)).to(dev) if not torch.is_tensor(sv_id) else sv_id, d, e, pn,
                         c.to(torch.float32) + idx * 1000.0))
    t1.record()
    if mine_out:
        packed = [torch.cat([m[j] for m in mine_out]) for j in range(5)]
    else:
        z = lambda *shape: torch.zeros(shape, dtype=torch.float64, device=dev)      # noqa: E731
        packed = [z(0), z(0), z(0), z(0), z(0, 3)]
    out = gather_region_scores(*packed, n_regions_total, dev, local_timings)
    t2.record()
    torch.cuda.synchronize(dev)
    timings.update(local_timings)
    timings.update(infer_and_score_ms=t0.elapsed_time(t1), scoring_ms=float(sum(a.elapsed_time(b) for a, b in score_ms)),
                   gather_ms=t1.elapsed_time(t2), device_total_ms=t0.elapsed_time(t2))
    flags = None
    if select_with is not None:
        h0 = time.perf_counter()
        flags = score.select_regions(select_with[0], out[0], out[1], out[2], out[3], select_with[1], device=dev, timings=timings)
        timings["selection_ms"] = (time.perf_counter() - h0) * 1e3
    return out + (flags, timings)


def infer_and_score_sequence(engine, raws, xyz, regions, seed=0, inf_reps=8, nei_num=24, dis_thresh=0.1, device="cuda"):
    """The whole LiDAL chain for one sequence on ONE GPU with caller-registered coordinates (kept for callers that hold
    KD-tree pickles): raw scan -> 8 TTA views -> network -> tail -> resident prob map -> inter-frame scoring -> region means.
    raws: per-frame float32 [Np,4] sensor-frame points (host or device); xyz: per-frame float64 [Np,3] registered coordinates;
    regions: per-frame (sv_id, sv2point).  Returns (per-frame worker_func tuples, timings dict in ms)."""
    from . import score
    dev = torch.device(device)
    ev = lambda: torch.cuda.Event(enable_timing=True)          # noqa: E731
    t0, t1, t2 = ev(), ev(), ev()
    scorer = score.SequenceScorer(dev, nei_num, dis_thresh)
    t0.record()
    for i, raw in enumerate(raws):
        raw_dev = torch.as_tensor(raw).to(dev)
        prob, _pred = prob_inference_frame(engine, raw_dev, seed + i, inf_reps)
        scorer.add_frame(xyz[i], prob, *regions[i])
    t1.record()
    out = [scorer.score_frame(i) for i in range(len(raws))]
    t2.record()
    torch.cuda.synchronize()
    n = len(raws)
    return out, {"prob_inference_ms_per_frame": t0.elapsed_time(t1) / n, "scoring_ms_per_frame": t1.elapsed_time(t2) / n,
                 "frames_per_sec": 1e3 * n / t0.elapsed_time(t2), "frames": n}
