"""CPU, world_size 2 over gloo: frame sharding, neighbour-window exchange and the single all_gather of region scores
reproduce the single-process result exactly (scorer = the pinned CPU oracle)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _case():
    from lidal_b200 import synth
    seq = synth.make_sequence(30, "NU", seed=4, max_points=400, step=0.15)
    probs = [synth.synthetic_probs(seq.xyz[i], 16, 70 + i) for i in range(seq.n_frames)]
    return seq, probs


def _oracle_score_frame(cache):
    import lidal_scoring as orc

    def score(fid, held, sv_id, sv2point):
        n = max(held) + 1
        nids = orc.neighbour_ids(fid, cache["n_frames"])
        for f in [fid] + nids:
            if f not in cache:
                cache[f] = orc.build_trees([held[f][0].numpy()])[0]
        d, e, _ = orc.score_points(held[fid][1].numpy(), held[fid][0].numpy(), [held[f][1].numpy() for f in nids],
                                   [cache[f] for f in nids])
        return orc.reduce_regions(d, e, held[fid][0].numpy(), sv_id, sv2point)
    return score


def _worker(rank, world, port, out_dir):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lidal_b200 import pipeline
    seq, probs = _case()
    n = seq.n_frames
    own = pipeline.frame_shard(n, world, rank)
    frames = {f: (torch.from_numpy(seq.xyz[f]), torch.from_numpy(probs[f])) for f in own}
    regions = {f: (seq.sv_id[f], seq.sv2point[f]) for f in own}
    n_points = [x.shape[0] for x in seq.xyz]
    out = pipeline.score_sequence_sharded(frames, n_points, 16, n, regions, _oracle_score_frame({"n_frames": n}),
                                          n_regions_total=int(seq.sv_id[-1][-1]) + 1)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), d=out[0], e=out[1], n=out[2], c=out[3])
    dist.barrier()
    dist.destroy_process_group()


def test_frame_shard_and_needs():
    from lidal_b200 import pipeline
    assert list(pipeline.frame_shard(10, 4, 0)) == [0, 1, 2] and list(pipeline.frame_shard(10, 4, 3)) == [9]
    assert [pipeline.owner_of(f, 10, 4) for f in range(10)] == [0, 0, 0, 1, 1, 1, 2, 2, 2, 3]
    covered = sorted(f for r in range(4) for f in pipeline.frame_shard(1000, 4, r))
    assert covered == list(range(1000))
    need = pipeline.needed_frames(pipeline.frame_shard(1000, 8, 3), 1000)
    assert need == list(range(375 - 12, 500 + 12))                       # interior rank: a 12-frame halo each side
    assert pipeline.needed_frames(range(0, 5), 1000) == list(range(0, 25))  # sequence start: reflected window


def test_frame_shard_equals_the_reference_data_loader(golden):
    """tests/golden/shard.npz: the frames the reference's own SK_Dataloader(gpu_num, gpu_rank).score_data_loader hands to each
    rank (dataset/sk_dataloader.py:196-198, run unmodified in the build container)."""
    from lidal_b200 import pipeline
    rows = golden["shard"]["rows"]
    assert len(rows) == 38
    for n, world, rank, start, stop in rows.tolist():
        own = pipeline.frame_shard(n, world, rank)
        assert (own.start, own.stop) == (start, stop), (n, world, rank)
        assert all(pipeline.owner_of(f, n, world) == rank for f in own)


@pytest.mark.timeout(300)
def test_two_rank_sharded_scoring_equals_single_process(tmp_path):
    sys.path[:0] = [os.path.join(ROOT, "oracle")]
    import lidal_scoring as orc
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    seq, probs = _case()
    trees = orc.build_trees(seq.xyz)
    n_regions = int(seq.sv_id[-1][-1]) + 1
    d, e, pn, c = np.zeros(n_regions, np.float32), np.zeros(n_regions, np.float32), np.zeros(n_regions, int), np.zeros((n_regions, 3), np.float32)
    for f in range(seq.n_frames):
        sv_id, sd, se, sn, sc = orc.score_frame(f, probs, seq.xyz, trees, seq.sv_id[f], seq.sv2point[f])
        d[sv_id], e[sv_id], pn[sv_id], c[sv_id] = sd, se, sn, sc
    for r in range(2):
        got = np.load(tmp_path / f"rank{r}.npz")
        assert np.array_equal(got["d"], d) and np.array_equal(got["e"], e)
        assert np.array_equal(got["n"], pn) and np.array_equal(got["c"], c)


# ------------------------------------------------------------------ many-sequence datasets: whole sequences per rank
SEQ_LENGTHS = [27, 25, 30, 26]          # the reference's window rule needs >= 25 frames per sequence (LiDAL.py:41-42)


def _dataset_case():
    """Four short NU-shaped sequences with dataset-wide region ids (prepare_supervoxel_kmeans_sk.py:67-69)."""
    from lidal_b200 import synth
    seqs, nxt = [], 0
    for s, n in enumerate(SEQ_LENGTHS):
        seq = synth.make_sequence(n, "NU", seed=20 + s, sv_id_start=nxt, max_points=200, step=0.15)
        nxt = int(seq.sv_id[-1][-1]) + 1
        probs = [synth.synthetic_probs(seq.xyz[i], 16, 1000 * s + i) for i in range(n)]
        seqs.append((seq, probs))
    return seqs, nxt


def _oracle_score_sequence(seqs):
    import lidal_scoring as orc

    def score(idx):
        seq, probs = seqs[idx]
        trees = orc.build_trees(seq.xyz)
        outs = [orc.score_frame(f, probs, seq.xyz, trees, seq.sv_id[f], seq.sv2point[f]) for f in range(seq.n_frames)]
        return tuple(np.concatenate([o[j] for o in outs]) for j in range(5))
    return score


def _dataset_worker(rank, world, port, out_dir):
    sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle")]
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lidal_b200 import pipeline
    seqs, n_regions = _dataset_case()
    out = pipeline.score_sequences_sharded(SEQ_LENGTHS, _oracle_score_sequence(seqs), n_regions)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), d=out[0], e=out[1], n=out[2], c=out[3])
    dist.barrier()
    dist.destroy_process_group()


def test_pack_sequences():
    from lidal_b200 import pipeline
    assert pipeline.pack_sequences([7, 5, 9, 4], 2) == [[2, 3], [0, 1]]            # 9 -> r0, 7 -> r1, 5 -> r1 (7 < 9), 4 -> r0: loads 13 / 12
    assert pipeline.pack_sequences([3, 3, 3], 1) == [[0, 1, 2]]
    assert pipeline.pack_sequences([], 4) == [[], [], [], []]
    assert pipeline.pack_sequences([5, 1], 4) == [[0], [1], [], []]                # more ranks than sequences: idle ranks
    rng = np.random.default_rng(0)
    counts = rng.integers(30, 50, 850).tolist()                                     # nuScenes-shaped: 850 scenes of ~40 frames
    bins = pipeline.pack_sequences(counts, 8)
    assert sorted(i for b in bins for i in b) == list(range(850))                   # every sequence exactly once
    loads = [sum(counts[i] for i in b) for b in bins]
    assert max(loads) - min(loads) <= max(counts)                                   # LPT: within one sequence of balanced
    assert max(loads) <= 1.01 * sum(counts) / 8
    assert all(b == sorted(b) for b in bins)
    assert pipeline.sequence_region_offsets([4, 0, 6]) == [0, 4, 4]


@pytest.mark.timeout(300)
def test_two_rank_sequence_packing_equals_reference_loop(tmp_path):
    """LiDAL.py:185-218 over several sequences (centres offset by idx * 1000.0) == the 2-rank whole-sequence sharding."""
    sys.path[:0] = [os.path.join(ROOT, "oracle")]
    mp.spawn(_dataset_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    seqs, n_regions = _dataset_case()
    score = _oracle_score_sequence(seqs)
    d, e = np.zeros(n_regions, np.float32), np.zeros(n_regions, np.float32)
    pn, c = np.zeros(n_regions, int), np.zeros((n_regions, 3), np.float32)
    for idx in range(len(seqs)):                                                    # the reference's loop, one process
        sv_id, sd, se, sn, sc = score(idx)
        d[sv_id], e[sv_id], pn[sv_id] = sd, se, sn
        c[sv_id] = sc + idx * 1000.0                                                # LiDAL.py:218
    assert c[:, 0].max() > 2900.0
    for r in range(2):
        got = np.load(tmp_path / f"rank{r}.npz")
        assert np.array_equal(got["d"], d) and np.array_equal(got["e"], e)
        assert np.array_equal(got["n"], pn) and np.array_equal(got["c"], c)
    from lidal_b200 import pipeline                                                  # world size 1 (no process group): same arrays
    one = pipeline.score_sequences_sharded(SEQ_LENGTHS, score, n_regions)
    assert np.array_equal(one[0], d) and np.array_equal(one[3], c)
