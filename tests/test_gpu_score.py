"""GPU: LiDAL inter-frame scoring and selection vs the reference-pinned goldens and the oracle.

Bars (BASELINE.json north_star): matches / region ids / selected flags bit-exact; scores within 1e-5 relative."""
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))


@pytest.fixture(scope="module")
def scorer_and_inputs():
    from make_golden import build_scoring_inputs, scoring_case
    from lidal_b200 import score
    seq, probs = build_scoring_inputs(scoring_case())
    score.init_worker(False, 24, 0.1, "00", probs, seq.xyz, list(zip(seq.sv_id, seq.sv2point)))
    return score, seq, probs


def test_worker_func_vs_reference_golden(scorer_and_inputs, golden):
    score, seq, probs = scorer_and_inputs
    g = golden["scoring"]
    for i in range(seq.n_frames):
        sv_id, d, e, pn, c = score.worker_func(i)
        assert sv_id.dtype == np.int64 and d.dtype == np.float32 and e.dtype == np.float32 and c.dtype == np.float32
        assert np.array_equal(sv_id, g["sv_id"][i]) and np.array_equal(pn, g["sv_pnums"][i])
        np.testing.assert_allclose(d, g["sv_interds"][i], rtol=1e-5, atol=1e-9)
        np.testing.assert_allclose(e, g["sv_interes"][i], rtol=1e-5)
        np.testing.assert_allclose(c, g["sv_centers"][i], rtol=1e-6, atol=1e-6)


def test_per_point_scores_and_matches_exact(scorer_and_inputs, golden):
    """Per-point divergence / entropy / match counts for three frames (first, middle, last: both reflection rules)."""
    import lidal_scoring as orc
    score, seq, probs = scorer_and_inputs
    g = golden["scoring"]
    scorer = score.var_dict["scorer"]
    trees = orc.build_trees(seq.xyz)
    for j, fid in enumerate(g["pt_frames"]):
        d, e, cnt, nn = scorer.score_points(int(fid), want_nn=True)
        assert np.array_equal(cnt.cpu().numpy(), g[f"pt_c{j}"].astype(np.int32))          # bit-exact match counts
        np.testing.assert_allclose(d.cpu().numpy(), g[f"pt_d{j}"], rtol=1e-5, atol=1e-9)
        np.testing.assert_allclose(e.cpu().numpy(), g[f"pt_e{j}"], rtol=1e-5)
        # nearest-neighbour indices identical to the KD-tree's
        for col, n in enumerate(orc.neighbour_ids(int(fid), seq.n_frames)):
            dist, idx = trees[n].query(seq.xyz[int(fid)], k=1)
            want = np.where(dist[:, 0] <= 0.1, idx[:, 0], -1)
            assert np.array_equal(nn[:, col].cpu().numpy(), want), (fid, n)
        frac_exact = float((d.cpu().numpy() == g[f"pt_d{j}"]).mean())
        print(f"frame {fid}: {frac_exact:.4f} of per-point divergences bit-identical")


def test_sequence_properties_full_size():
    """SK-sized frames (too slow for the KD-tree oracle in a unit test): size-independent properties."""
    from lidal_b200 import score, synth
    seq = synth.make_sequence(25, "SK", seed=3)
    probs = [synth.synthetic_probs(seq.xyz[i], 19, 900 + i) for i in range(25)]
    sc = score.SequenceScorer()
    for i in range(25):
        sc.add_frame(seq.xyz[i], probs[i], seq.sv_id[i], seq.sv2point[i])
    d, e, cnt = sc.score_points(12)
    d, e, cnt = d.cpu().numpy(), e.cpu().numpy(), cnt.cpu().numpy()
    assert d.shape[0] == seq.xyz[12].shape[0] and (d >= -1e-6).all() and np.isfinite(d).all()
    assert (e >= 0).all() and (e <= np.log(19) + 1e-5).all()
    assert (cnt >= 0).all() and (cnt <= 24).all() and cnt.mean() > 1
    assert (d[cnt == 0] == 0).all()
    # identical neighbour => zero divergence, entropy of the query distribution itself
    sc2 = score.SequenceScorer()
    for i in range(25):
        sc2.add_frame(seq.xyz[0], probs[0], seq.sv_id[0], seq.sv2point[0])
    d2, e2, c2 = sc2.score_points(5)
    assert (c2.cpu().numpy() == 24).all() and float(d2.abs().max()) < 1e-6
    p = probs[0].astype(np.float64)
    np.testing.assert_allclose(e2.cpu().numpy(), -(p * np.log(p)).sum(1), rtol=1e-4, atol=1e-5)
    sv_id, sd, se, pn, c = sc.score_frame(12)
    assert pn.sum() == seq.xyz[12].shape[0] and len(sv_id) == 20


def test_selection_bit_exact(golden):
    from lidal_b200 import score
    g = golden["selection"]
    out = score.select_regions(g["sv_flags"].copy(), g["sv_interds"], g["sv_interes"], g["sv_pnums"], g["sv_centers"],
                               int(g["tight_tpn"]))
    assert np.array_equal(out, g["tight_out"])
    out = score.select_regions(g["sv_flags"].copy(), g["loose_interds"], g["sv_interes"], g["sv_pnums"],
                               g["sv_centers"], int(g["loose_tpn"]))
    assert np.array_equal(out, g["loose_out"])


def test_argsort_and_region_pairs_vs_numpy():
    from lidal_b200 import score
    rng = np.random.default_rng(1)
    keys = rng.normal(size=50000).astype(np.float32)
    keys[::7] = 0.0
    keys[5] = -0.0
    order = score.argsort_f32(torch.from_numpy(keys).cuda()).cpu().numpy()
    assert np.array_equal(keys[order], np.sort(keys))
    assert np.array_equal(order, np.argsort(keys, kind="stable"))
    c = (rng.random((4000, 3)) * [200, 30, 3]).astype(np.float32)
    ptr, idx = score.region_pairs(torch.from_numpy(c).cuda(), 5.0)
    ptr, idx = ptr.cpu().numpy(), idx.cpu().numpy()
    for i in rng.integers(0, 4000, 200):
        dist = np.sqrt(np.square(c[i] - c).sum(1, dtype=np.float32))
        want = np.where((dist < np.float32(5.0)) & (np.arange(4000) != i))[0]
        assert np.array_equal(np.sort(idx[ptr[i]:ptr[i + 1]]), want)


def test_frame_level_baselines_vs_reference_formulas():
    """F3: softmax entropy / margin / least confidence frame scores vs the reference's numpy expressions."""
    from scipy.stats import entropy
    from lidal_b200 import score, synth
    rng = np.random.default_rng(3)
    for n_cls in (19, 16):
        xyz = rng.random((50000, 3)) * 40
        prob = synth.synthetic_probs(xyz, n_cls, 7)
        prob[::97, 3] = prob[::97, 5]                      # exact top-2 ties
        ent, mar, conf = score.frame_level_scores(torch.from_numpy(prob).cuda())
        srt = np.sort(prob, axis=-1)
        np.testing.assert_allclose(ent, np.mean(entropy(prob, axis=1)), rtol=1e-5)
        np.testing.assert_allclose(mar, np.mean(srt[:, -1] - srt[:, -2]), rtol=1e-5)
        np.testing.assert_allclose(conf, np.mean(srt[:, -1]), rtol=1e-5)
