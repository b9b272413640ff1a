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


def _write_reference_files(d, seq, probs, sv_id_offset=0):
    """The reference's on-disk ABI between prob_inference / pre-processing and scoring: prob .npy (score/prob_inference.py:129),
    sklearn KDTree .pickle (dataset/prepare_kdtree_sk.py:83-88), (sv_id, sv2point) .pickle (prepare_supervoxel_kmeans_sk.py:60-74)."""
    import pickle
    from sklearn.neighbors import KDTree
    pf, kf, sf = [], [], []
    for i in range(seq.n_frames):
        pf.append(f"{d}/p{i:06d}.npy"); np.save(pf[-1], probs[i])
        kf.append(f"{d}/k{i:06d}.pickle")
        with open(kf[-1], "wb") as f:
            pickle.dump(KDTree(seq.xyz[i]), f)
        sf.append(f"{d}/s{i:06d}.pickle")
        with open(sf[-1], "wb") as f:
            pickle.dump((seq.sv_id[i] + sv_id_offset, seq.sv2point[i]), f)
    return pf, kf, sf


def test_init_worker_from_reference_files(golden, tmp_path):
    """B2 through the file-path branch: the same calls score/sv_level/LiDAL.py:204-206 makes, on files in the reference's formats."""
    from make_golden import build_scoring_inputs, scoring_case
    from lidal_b200 import score
    seq, probs = build_scoring_inputs(scoring_case())
    pf, kf, sf = _write_reference_files(str(tmp_path), seq, probs)
    g = golden["scoring"]
    for sv_pre in (False, True):
        score.init_worker(sv_pre, 24, 0.1, "00", pf, kf, sf)
        for i in (0, 7, 13, 25):
            out = score.worker_func(i)
            assert len(out) == (3 if sv_pre else 5)
            assert np.array_equal(out[0], g["sv_id"][i]) and out[0].dtype == np.int64
            np.testing.assert_allclose(out[1], g["sv_interds"][i], rtol=1e-5, atol=1e-9)
            np.testing.assert_allclose(out[2], g["sv_interes"][i], rtol=1e-5)
            if not sv_pre:
                assert np.array_equal(out[3], g["sv_pnums"][i])
                np.testing.assert_allclose(out[4], g["sv_centers"][i], rtol=1e-6, atol=1e-6)


def test_score_dataset_multi_sequence(golden, tmp_path):
    """LiDAL.py:185-222: per-sequence scoring scattered into the global arrays; centres of sequence idx shifted by idx * 1000
    (:218); with cached sv_pnums / sv_centers (sv_pre, :171-175) only the two score arrays are filled."""
    from make_golden import build_scoring_inputs, scoring_case
    from lidal_b200 import score
    seq, probs = build_scoring_inputs(scoring_case())
    g = golden["scoring"]
    n_seq_regions = int(g["sv_id"].max()) + 1
    (tmp_path / "a").mkdir(); (tmp_path / "b").mkdir()
    seq_a = _write_reference_files(str(tmp_path / "a"), seq, probs)
    seq_b = _write_reference_files(str(tmp_path / "b"), seq, probs, sv_id_offset=n_seq_regions)
    d, e, pn, c, sv_pre = score.score_dataset([seq_a, seq_b])
    assert not sv_pre and d.shape[0] == 2 * n_seq_regions and d.dtype == np.float32 and c.dtype == np.float32
    ids = g["sv_id"].reshape(-1)
    for k, off in enumerate((0, n_seq_regions)):
        np.testing.assert_allclose(d[ids + off], g["sv_interds"].reshape(-1), rtol=1e-5, atol=1e-9)
        np.testing.assert_allclose(e[ids + off], g["sv_interes"].reshape(-1), rtol=1e-5)
        assert np.array_equal(pn[ids + off], g["sv_pnums"].reshape(-1))
        np.testing.assert_allclose(c[ids + off], g["sv_centers"].reshape(-1, 3) + np.float32(k * 1000.0), rtol=1e-6, atol=1e-4)
    assert np.array_equal(d[:n_seq_regions], d[n_seq_regions:])                 # same inputs, same bits
    d2, e2, pn2, c2, sv_pre2 = score.score_dataset([seq_a, seq_b], sv_pnums=pn, sv_centers=c)
    assert sv_pre2 and pn2 is pn and c2 is c and np.array_equal(d2, d) and np.array_equal(e2, e)
