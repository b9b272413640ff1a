"""GPU: BASELINE-sized inputs (8 SemanticKITTI-shaped scans, ~0.75 M voxels) through size-independent properties -- the
CPU oracle would need minutes here, so the checks are invariants of the domain and agreement between independent paths."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sk_batch():
    from lidal_b200 import synth
    raw = synth.raycast_scan(123, "SK")
    coords, feats, inv = synth.tta_batch(raw, seed=3, inf_reps=8)          # the batch prob_inference sees: 8 views of one scan
    return raw, torch.from_numpy(coords).cuda(), torch.from_numpy(feats).cuda(), torch.from_numpy(inv).cuda()


def test_kernel_map_invariants_full_size(sk_batch):
    import lidal_b200.compat as ts
    from lidal_b200 import engine
    F = ts.nn.functional
    _, coords, _, _ = sk_batch
    n = coords.shape[0]
    assert n > 600_000
    km = F.build_kernel_map(coords, (1, 1, 1), (3, 3, 3), (1, 1, 1), (1, 1, 1))
    nbr = km.nbr
    assert torch.equal(nbr[13], torch.arange(n, dtype=torch.int, device="cuda"))           # centre offset = identity
    for k in (0, 5, 12):                                                                    # symmetry: k <-> 26 - k
        o = torch.nonzero(nbr[k] >= 0).squeeze(1)
        assert torch.equal(nbr[26 - k][nbr[k][o].long()].long(), o)
    assert int(km.nbsizes.sum()) == int((nbr >= 0).sum()) == km.nbmaps.shape[0]
    # neighbours really are the claimed offsets away, and never cross scans
    offs = ts.nn.utils.get_kernel_offsets(3, 1, 1, device="cuda")
    for k in (3, 22):
        o = torch.nonzero(nbr[k] >= 0).squeeze(1)[:100000]
        i = nbr[k][o].long()
        assert torch.equal(coords[i, :3], coords[o, :3] + offs[k]) and torch.equal(coords[i, 3], coords[o, 3])
    # engine maps: every fine voxel has exactly one parent, parents are unique, mask-sorted tables are permutations
    m = engine.Maps(coords)
    for lvl in range(4):
        up_sorted, perm = m.nbr_up[lvl]
        assert int((up_sorted >= 0).sum()) == m.n[lvl]
        assert torch.equal(torch.sort(perm.long()).values, torch.arange(m.n[lvl], device="cuda"))
        cn = m.coords[lvl + 1]
        key = (cn[:, 3].long() << 48) | (cn[:, 0].long() << 32) | (cn[:, 1].long() << 16) | cn[:, 2].long()
        assert torch.unique(key).numel() == cn.shape[0]
        assert m.n[lvl + 1] < m.n[lvl]


@pytest.mark.parametrize("name", ["minkunet", "spvcnn"])
def test_engine_vs_dropin_full_size(sk_batch, name):
    """Two independent CUDA paths (module-by-module fp32 boundary vs fused bf16 engine) agree at full size; the TTA tail
    yields a proper distribution and the argmax of the mean."""
    import lidal_b200.compat as ts
    from lidal_b200 import score
    from lidal_b200.engine import InferenceEngine
    from lidal_b200.network import MinkUNet, SPVCNN, seeded_state_dict
    raw, coords, feats, inv = sk_batch
    model = (MinkUNet if name == "minkunet" else SPVCNN)(19, ts)
    model.load_state_dict(seeded_state_dict(model.state_dict()), strict=True)
    model = model.cuda().eval()
    with torch.no_grad():
        ref = model(ts.SparseTensor(feats, coords))[0]
    got = InferenceEngine(model)(coords, feats)
    assert got.shape == ref.shape == (coords.shape[0], 19) and bool(torch.isfinite(got).all())
    err = float((got.double() - ref.double()).norm() / ref.double().norm())
    print(f"{name}: engine vs drop-in at {coords.shape[0]} voxels: rel-L2 {err:.3e}")
    assert err < 1e-2
    prob, pred = score.tta_tail(got, inv, 8)
    assert prob.shape == (raw.shape[0], 19)
    assert float((prob.sum(1) - 1).abs().max()) < 1e-5 and float(prob.min()) >= 0
    assert torch.equal(pred, prob.argmax(1))


def test_gpu_voxelizer_full_size_roundtrip(sk_batch):
    """F1 at full size: inverse indices reconstruct every point's voxel; voxels are sorted and unique per view."""
    from lidal_b200 import voxelizer
    raw = sk_batch[0]
    c, f, inv = voxelizer.tta_batch_gpu(torch.from_numpy(raw).cuda(), seed=9, inf_reps=8)
    assert inv.shape[0] == 8 * raw.shape[0] and int(inv.max()) == c.shape[0] - 1 and int(inv.min()) == 0
    key = (c[:, 3].long() << 39) | (c[:, 0].long() << 26) | (c[:, 1].long() << 13) | c[:, 2].long()
    assert bool((key[1:] > key[:-1]).all())                                  # lexicographic (b, x, y, z), unique
    assert bool((c[:, :3] >= 0).all()) and bool((c[:, :3] < 8192).all())
    assert torch.unique(inv).numel() == c.shape[0]                           # every voxel is hit by at least one point


# ------------------------------------------------------------------------------------------ oracle parity AT BASELINE size
@pytest.mark.parametrize("name", ["minkunet", "spvcnn"])
def test_engine_logits_vs_oracle_full_sk_scan(name, oracle_ts):
    """One full SemanticKITTI-shaped scan (~100k voxels, BASELINE configs[0] shape) through the fused engine vs the reference's
    network on the CPU oracle (a few seconds of host time): logits within 1e-2 relative, as north_star states for
    bf16 operands / f32 accumulate vs the fp32 reference."""
    import lidal_b200.compat as ts
    from lidal_b200 import synth
    from lidal_b200.engine import InferenceEngine
    from lidal_b200.network import MinkUNet, SPVCNN, seeded_state_dict
    raw = synth.raycast_scan(123, "SK")
    coords, feats, _ = synth.tta_batch(raw, seed=3, inf_reps=1)
    assert coords.shape[0] > 80_000
    cls = MinkUNet if name == "minkunet" else SPVCNN
    ref = cls(19, oracle_ts)
    sd = seeded_state_dict(ref.state_dict())
    ref.load_state_dict(sd); ref.eval()
    torch.set_num_threads(max(1, __import__("os").cpu_count() or 1))
    with torch.no_grad():
        want, want_feat = ref(oracle_ts.SparseTensor(torch.from_numpy(feats), torch.from_numpy(coords)))
    dev = cls(19, ts); dev.load_state_dict(sd); dev = dev.cuda().eval()
    got, got_feat = InferenceEngine(dev)(torch.from_numpy(coords).cuda(), torch.from_numpy(feats).cuda(), return_feat=True)
    err = float((got.cpu().double() - want.double()).norm() / want.double().norm())
    err_f = float((got_feat.float().cpu().double() - want_feat.double()).norm() / want_feat.double().norm())
    agree = float((got.cpu().argmax(1) == want.argmax(1)).double().mean())
    print(f"{name}: {coords.shape[0]} voxels, logits rel-L2 {err:.3e}, out_feat rel-L2 {err_f:.3e}, argmax agreement {agree:.4f}")
    assert err < 1e-2 and err_f < 1e-2 and agree > 0.98


def test_score_points_vs_oracle_full_sk_frame():
    """One SK-sized query frame (~130k points) against its 24-frame window vs oracle.lidal_scoring.score_points (the
    reference's KD-tree loop, LiDAL.py:59-81; about a minute of host time): match counts bit-exact, D / H within 1e-5."""
    import lidal_scoring as orc
    from lidal_b200 import score, synth
    n_frames, fid = 25, 12
    seq = synth.make_sequence(n_frames, "SK", seed=3)
    probs = [synth.synthetic_probs(seq.xyz[i], 19, 900 + i) for i in range(n_frames)]
    sc = score.SequenceScorer()
    for i in range(n_frames):
        sc.add_frame(seq.xyz[i], probs[i], seq.sv_id[i], seq.sv2point[i])
    d, e, cnt = sc.score_points(fid)
    nids = orc.neighbour_ids(fid, n_frames)
    trees = orc.build_trees([seq.xyz[n] for n in nids])
    want_d, want_e, want_c = orc.score_points(probs[fid], seq.xyz[fid], [probs[n] for n in nids], trees)
    assert seq.xyz[fid].shape[0] > 120_000 and want_c.sum() > 1_000_000
    assert np.array_equal(cnt.cpu().numpy(), want_c.astype(np.int32))
    np.testing.assert_allclose(d.cpu().numpy(), want_d, rtol=1e-5, atol=1e-9)
    np.testing.assert_allclose(e.cpu().numpy(), want_e, rtol=1e-5)
    sv = sc.score_frame(fid)
    want_sv = orc.reduce_regions(want_d, want_e, seq.xyz[fid], seq.sv_id[fid], seq.sv2point[fid])
    assert np.array_equal(sv[0], want_sv[0]) and np.array_equal(sv[3], want_sv[3])
    np.testing.assert_allclose(sv[1], want_sv[1], rtol=1e-5)
    np.testing.assert_allclose(sv[2], want_sv[2], rtol=1e-5)


def test_select_regions_vs_oracle_20k_regions():
    """The global selection at config-3 size (1000 frames x 20 regions, budget of LiDAL.py:130) == oracle, both index paths."""
    import lidal_scoring as orc
    from lidal_b200 import score, synth
    flags, d, e, pn, c = synth.region_table(1000, seed=1)
    tpn = 2349559532
    want = orc.select_regions(flags.astype(np.float64), d, e, pn, c, tpn)
    for method in ("csr", "grid"):
        got = score.select_regions(flags.copy(), d, e, pn, c, tpn, method=method)
        assert np.array_equal(got, want), method
    assert (want == 1).sum() > 500 and (want == 2).sum() > 500
