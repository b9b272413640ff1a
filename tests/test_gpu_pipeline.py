"""The single-GPU legs of the sharded LiDAL drivers (lidal_b200.pipeline): `run_sequence_sharded` (BASELINE configs[2]) and
`run_dataset_sharded` (configs[4], whole sequences per rank) against a step-by-step composition of the same public pieces
-- prob_inference per frame, SequenceScorer, the reference's per-sequence loop with its `idx * 1000.0` centre offset
(score/sv_level/LiDAL.py:185-218).  The N > 1 legs of both drivers are covered on CPU over gloo (tests/test_dist_cpu.py)
and by bench.py --gpus N on the box."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

N_CLS = 16
# Two runs of SPVCNN on the same frame agree to rounding, not bit for bit: point_to_voxel sums a voxel's points in the order a
# counting sort with an atomic cursor lists them (csrc/frame.cu, segment_order) -- like the reference's atomicAdd voxelize,
# "neither order is canonical" -- so a 16-bit mean can flip its last bit and move a region score by ~1 ulp.
SCORE_TOL = dict(rtol=2e-3, atol=1e-7)
SEQ_LENGTHS = [26, 25]            # the reference's window rule needs >= 25 frames per sequence (LiDAL.py:41-42)
REGIONS = 20


@pytest.fixture(scope="module")
def engine():
    import lidal_b200.compat as ts
    from lidal_b200.engine import InferenceEngine
    from lidal_b200.network import SPVCNN, seeded_state_dict
    model = SPVCNN(N_CLS, ts)
    model.load_state_dict(seeded_state_dict(model.state_dict()), strict=True)
    return InferenceEngine(model.cuda().eval())


def _sequences(dev):
    from lidal_b200 import synth
    from lidal_b200.pipeline import sequence_region_offsets
    offs = sequence_region_offsets([n * REGIONS for n in SEQ_LENGTHS])
    return [synth.GpuSequence(n, "NU", seed=31 + s, device=dev, n_regions=REGIONS, sv_id_start=offs[s]) for s, n in enumerate(SEQ_LENGTHS)]


def _compose(engine, seq, seed, dev):
    """One sequence, frame by frame through the public pieces (no StreamPipeline, no side stream)."""
    from lidal_b200 import pipeline, score
    scorer = score.SequenceScorer(dev, n_total=seq.n_frames)
    for fid in range(seq.n_frames):
        raw, pose, sv_id, regions = seq.frame(fid)
        prob, _pred = pipeline.prob_inference_frame(engine, raw, seed + fid)
        scorer.add_frame(score.register_points(raw, pose), prob, sv_id, regions, fid=fid)
    return [scorer.score_frame(f) for f in range(seq.n_frames)]


def test_sequence_driver_equals_composition(engine):
    from lidal_b200 import pipeline
    dev = torch.device("cuda", torch.cuda.current_device())
    seq = _sequences(dev)[0]
    n_regions = seq.n_frames * REGIONS
    d, e, pn, c, flags, tm = pipeline.run_sequence_sharded(engine, seq.frame, seq.n_frames, N_CLS, n_regions, seed=5, device=dev)
    assert flags is None and tm["frames_own"] == seq.n_frames
    for sv_id, sd, se, sn, sc in _compose(engine, seq, 5, dev):
        np.testing.assert_allclose(d[sv_id], sd, **SCORE_TOL)
        np.testing.assert_allclose(e[sv_id], se, **SCORE_TOL)
        assert np.array_equal(pn[sv_id], sn) and np.array_equal(c[sv_id], sc)     # integer counts / float64 registration: exact
    assert np.isfinite(d).all() and (e > 0).any() and int(pn.sum()) > 0


def test_dataset_driver_equals_reference_loop(engine):
    """Whole-sequence sharding (one rank here): per-region arrays == LiDAL.py:185-218 restated over the composition."""
    from lidal_b200 import pipeline
    dev = torch.device("cuda", torch.cuda.current_device())
    seqs = _sequences(dev)
    n_regions = sum(SEQ_LENGTHS) * REGIONS
    flags0 = np.zeros(n_regions, int)
    flags0[: REGIONS] = 1
    d, e, pn, c, flags, tm = pipeline.run_dataset_sharded(engine, lambda idx: seqs[idx].frame, SEQ_LENGTHS, N_CLS, n_regions, seed=5,
                                                          device=dev, select_with=(flags0, 40 * int(3e4) * 100))
    assert tm["sequences_own"] == 2 and tm["frames_own"] == sum(SEQ_LENGTHS)
    D, E = np.zeros(n_regions, np.float32), np.zeros(n_regions, np.float32)
    PN, C = np.zeros(n_regions, int), np.zeros((n_regions, 3), np.float32)
    for idx, seq in enumerate(seqs):
        for sv_id, sd, se, sn, sc in _compose(engine, seq, 5 + idx * 100003, dev):
            D[sv_id], E[sv_id], PN[sv_id] = sd, se, sn
            C[sv_id] = sc + idx * 1000.0                                          # LiDAL.py:218
    np.testing.assert_allclose(d, D, **SCORE_TOL)
    np.testing.assert_allclose(e, E, **SCORE_TOL)
    assert np.array_equal(pn, PN) and np.array_equal(c, C)
    assert C[SEQ_LENGTHS[0] * REGIONS:, 0].min() > 900.0
    # the replicated selection on the driver's own arrays: flags bit-equal to the CPU restatement of LiDAL.py:230-325
    import lidal_scoring as orc
    want = orc.select_regions(flags0, d, e, pn, c, 40 * int(3e4) * 100)
    assert np.array_equal(flags, want) and (flags == 1).sum() > REGIONS
