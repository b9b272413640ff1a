"""GPU: the per-frame device steps either side of the network (csrc/frame.cu) vs the oracle restatements that are pinned
to the reference's own code (oracle/lidal_extra.py, tests/golden/{voxelizer,extra}.npz).

Bars: integer / index outputs and float64 registration bit-exact; float32 region means bit-exact (numpy's summation order
is reproduced); log-based scores within 1e-5 relative (libm differences only)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_register_points_bit_equal(golden):
    import lidal_extra as ox
    from lidal_b200 import score, synth
    g = golden["extra"]
    seq = synth.make_sequence(2, "NU", seed=int(g["seq_seed"]), max_points=int(g["max_points"]))
    raw, pose = seq.raw[int(g["frame"])], g["pose"]
    xyz = score.register_points(torch.from_numpy(raw).cuda(), pose).cpu().numpy()
    assert xyz.dtype == np.float64 and np.array_equal(xyz, ox.register_points(raw, pose))
    assert np.array_equal(xyz[:64], g["xyz_head"])
    # a full-size scan under an arbitrary (non-orthonormal, large translation) pose: still bit-equal
    raw = synth.raycast_scan(5, "SK")
    pose = np.random.default_rng(1).normal(0, 1, (4, 4)); pose[:3, 3] *= 1000.0; pose[3] = [0, 0, 0, 1]
    xyz = score.register_points(torch.from_numpy(raw).cuda(), pose).cpu().numpy()
    assert np.array_equal(xyz, ox.register_points(raw, pose))


@pytest.mark.parametrize("kind,stride,seed", [("NU", 3, 3), ("SK", 1, 11)])
def test_batched_tta_voxelizer_bit_exact(kind, stride, seed, golden):
    """All 8 views in two launches + one unique == the reference's SK_Dataset('score') + collate_fn (golden for the NU case,
    oracle.lidal_extra.score_batch for the full-size SK scan); the per-view device path agrees too."""
    import lidal_extra as ox
    from lidal_b200 import synth, voxelizer
    raw = synth.raycast_scan(21 if kind == "NU" else 9, kind)[::stride]
    want_c, want_f, want_i = ox.score_batch(raw, seed, 8)
    raw_dev = torch.from_numpy(raw).cuda()
    for fn in (voxelizer.tta_batch_gpu, voxelizer.tta_batch_gpu_per_view):
        c, f, inv = fn(raw_dev, seed, 8)
        assert c.dtype == torch.int32 and f.dtype == torch.float32 and inv.dtype == torch.int64
        assert np.array_equal(c.cpu().numpy(), want_c), fn.__name__
        assert np.array_equal(f.cpu().numpy(), want_f), fn.__name__
        assert np.array_equal(inv.cpu().numpy(), want_i), fn.__name__
    if kind == "NU":
        g = golden["voxelizer"]
        assert c.shape[0] == int(g["n_vox"]) and np.array_equal(c[:256].cpu().numpy(), g["coords_head"])


def test_tta_tail_out_feat_mean():
    import lidal_extra as ox
    from lidal_b200 import score
    rng = np.random.default_rng(2)
    nv, npts, reps, c = 5000, 1700, 8, 96
    feat = torch.from_numpy(rng.normal(0, 1, (nv, c)).astype(np.float32)).cuda().bfloat16()
    logits = torch.from_numpy(rng.normal(0, 2, (nv, 19)).astype(np.float32)).cuda()
    inv = rng.integers(0, nv, reps * npts)
    prob, pred, mean = score.tta_tail(logits, torch.from_numpy(inv).cuda(), reps, out_feat=feat)
    want = ox.outfeat_mean(feat.float().cpu().numpy(), inv, reps)
    assert mean.dtype == torch.float32 and np.array_equal(mean.cpu().numpy(), want)        # float32 adds in view order: exact
    prob2, pred2 = score.tta_tail(logits, torch.from_numpy(inv).cuda(), reps)
    assert torch.equal(prob, prob2) and torch.equal(pred, pred2)


def test_segment_entropy_and_redal_vs_reference_golden(golden):
    import lidal_extra as ox
    from test_oracle_extra import extra_inputs
    from lidal_b200 import score
    g = golden["extra"]
    seq, f, prob, outfeat = extra_inputs(g)
    n_cls = int(g["n_cls"])
    pred = np.argmax(prob, 1)
    got = score.segment_entropy(torch.from_numpy(pred).cuda(), seq.sv2point[f], n_cls)
    assert abs(got - float(g["segment_entropy"])) <= 1e-12 * abs(float(g["segment_entropy"]))
    out = score.redal_region_scores(torch.from_numpy(prob).cuda(), torch.from_numpy(outfeat).cuda(), g["curvature"].astype(np.float32),
                                    seq.sv_id[f], seq.sv2point[f])
    assert np.array_equal(out[0], seq.sv_id[f]) and np.array_equal(out[3], g["redal_pnums"])
    np.testing.assert_allclose(out[1], g["redal_scores"], rtol=1e-5)
    assert out[2].dtype == np.float32 and np.array_equal(out[2], g["redal_feats"])        # sequential float32 column sums: exact
    # full-size frame: 20 regions of ~6.6k points
    raw = __import__("lidal_b200").synth.raycast_scan(2, "SK")
    _, sv2p = __import__("lidal_b200").synth.balanced_regions(raw)
    rng = np.random.default_rng(5)
    big_pred = rng.integers(0, 19, raw.shape[0])
    want = ox.segment_entropy(big_pred, sv2p, 19)
    got = score.segment_entropy(torch.from_numpy(big_pred).cuda(), sv2p, 19)
    assert abs(got - want) <= 1e-12 * abs(want)


@pytest.mark.parametrize("sizes", [[1, 3, 7, 8, 9, 64, 127, 128, 129, 200, 1000], [6575] * 20, [40000, 131, 70001]])
def test_region_means_follow_numpy_summation_order(sizes):
    """sv_interes = intere[p_ids].mean() (float32 pairwise) and sv_interds = interd[p_ids].mean() (float64 pairwise) are
    reproduced BIT FOR BIT for every region size class of numpy's pairwise_sum (n < 8, n <= 128, recursive splits)."""
    from lidal_b200 import _lib as L
    rng = np.random.default_rng(len(sizes))
    n = int(sum(sizes)) + 100
    perm = rng.permutation(n)
    sv2p, o = [], 0
    for s in sizes:
        sv2p.append(np.sort(perm[o:o + s])); o += s
    interd = rng.gamma(2.0, 0.05, n)
    intere = rng.uniform(0, 2.9, n).astype(np.float32)
    xyz = rng.normal(0, 30, (n, 3))
    from lidal_b200.score import regions_to_csr
    ptr, pts = regions_to_csr(sv2p, "cuda")
    r = len(sizes)
    sv_d = torch.empty(r, dtype=torch.float32, device="cuda"); sv_e = torch.empty(r, dtype=torch.float32, device="cuda")
    sv_n = torch.empty(r, dtype=torch.int64, device="cuda"); sv_c = torch.empty((r, 3), dtype=torch.float32, device="cuda")
    d_dev, e_dev, x_dev = torch.from_numpy(interd).cuda(), torch.from_numpy(intere).cuda(), torch.from_numpy(xyz).cuda()
    L.check(L.lib().lb_region_reduce(L.ptr(d_dev), L.ptr(e_dev), L.ptr(x_dev), L.ptr(ptr), L.ptr(pts), r, L.ptr(sv_d), L.ptr(sv_e),
                                     L.ptr(sv_n), L.ptr(sv_c), L.stream()))
    want_d = np.array([interd[p].mean() for p in sv2p]).astype(np.float32)
    want_e = np.array([intere[p].mean() for p in sv2p], dtype=np.float32)
    want_c = np.stack([xyz[p].mean(0) for p in sv2p]).astype(np.float32)
    assert np.array_equal(sv_e.cpu().numpy(), want_e)
    assert np.array_equal(sv_d.cpu().numpy(), want_d)
    assert np.array_equal(sv_n.cpu().numpy(), np.array(sizes))
    np.testing.assert_allclose(sv_c.cpu().numpy(), want_c, rtol=1e-6, atol=1e-6)
    # lb_region_mean_f32 is the same float32 mean on its own
    out = torch.empty(r, dtype=torch.float32, device="cuda")
    L.check(L.lib().lb_region_mean_f32(L.ptr(e_dev), L.ptr(ptr), L.ptr(pts), r, L.ptr(out), L.stream()))
    assert np.array_equal(out.cpu().numpy(), want_e)


def test_engine_out_feat_vs_oracle(small_scan, oracle_ts):
    """The model's second output (96-channel features) from the fused engine vs the reference network on the oracle."""
    import lidal_b200.compat as ts
    from lidal_b200.engine import InferenceEngine
    from lidal_b200.network import MinkUNet, SPVCNN, seeded_state_dict
    coords, feats, _ = small_scan
    for cls, ncls in ((MinkUNet, 19), (SPVCNN, 16)):
        ref = cls(ncls, oracle_ts)
        sd = seeded_state_dict(ref.state_dict())
        ref.load_state_dict(sd); ref.eval()
        with torch.no_grad():
            _, want = ref(oracle_ts.SparseTensor(torch.from_numpy(feats), torch.from_numpy(coords)))
        dev = cls(ncls, ts); dev.load_state_dict(sd); dev = dev.cuda().eval()
        logits, feat = InferenceEngine(dev)(torch.from_numpy(coords).cuda(), torch.from_numpy(feats).cuda(), return_feat=True)
        assert feat.shape == (coords.shape[0], 96)
        err = float((feat.float().cpu().double() - want.double()).norm() / want.double().norm())
        assert err < 1e-2, (cls.__name__, err)
