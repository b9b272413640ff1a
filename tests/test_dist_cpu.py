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
