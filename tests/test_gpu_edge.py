"""GPU: edge cases -- empty inputs, tiny inputs (< one 128-row tile), a ragged batch with an empty scan, isolated voxels."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _models(ncls=19):
    import lidal_b200.compat as ts
    from lidal_b200.network import MinkUNet, SPVCNN, seeded_state_dict
    out = []
    for cls in (MinkUNet, SPVCNN):
        m = cls(ncls, ts)
        m.load_state_dict(seeded_state_dict(m.state_dict()), strict=True)
        out.append(m.cuda().eval())
    return ts, out


def _ref_logits(oracle_ts, cls_name, coords, feats, ncls=19):
    from lidal_b200 import network
    m = getattr(network, cls_name)(ncls, oracle_ts)
    m.load_state_dict(network.seeded_state_dict(m.state_dict()), strict=True)
    m.eval()
    with torch.no_grad():
        return m(oracle_ts.SparseTensor(torch.from_numpy(feats), torch.from_numpy(coords)))[0]


@pytest.mark.parametrize("n", [0, 1, 5, 127, 129])
def test_tiny_and_empty_inputs(oracle_ts, n):
    from lidal_b200.engine import InferenceEngine
    ts, models = _models()
    rng = np.random.default_rng(n)
    coords = np.unique(np.concatenate([rng.integers(100, 110, (n, 3)), np.zeros((n, 1), int)], 1), axis=0).astype(np.int32)
    feats = rng.normal(size=(coords.shape[0], 4)).astype(np.float32)
    c, f = torch.from_numpy(coords).cuda(), torch.from_numpy(feats).cuda()
    for model, name in zip(models, ("MinkUNet", "SPVCNN")):
        with torch.no_grad():
            a = model(ts.SparseTensor(f, c))[0]
        b = InferenceEngine(model)(c, f)
        assert a.shape == b.shape == (coords.shape[0], 19)
        if coords.shape[0]:
            want = _ref_logits(oracle_ts, name, coords, feats)
            for got in (a, b):
                err = float((got.cpu().double() - want.double()).norm() / want.double().norm().clamp_min(1e-9))
                assert err < 2e-2, (name, n, err)


def test_ragged_batch_with_empty_scan(oracle_ts):
    """Batch ids {0, 2}: scan 1 contributed no voxels (dataset/sk_dataset.py:207-209 would still number it)."""
    from lidal_b200 import synth
    from lidal_b200.engine import InferenceEngine
    ts, models = _models()
    raw = synth.raycast_scan(7, "NU")[::16]
    rs = np.random.RandomState(1)
    c0, f0, _ = synth.score_transform(raw, rs)
    c2, f2, _ = synth.score_transform(raw[::2], rs)
    coords = np.concatenate([np.c_[c0, np.zeros(len(c0), int)], np.c_[c2, np.full(len(c2), 2)]]).astype(np.int32)
    feats = np.concatenate([f0, f2]).astype(np.float32)
    want = _ref_logits(oracle_ts, "MinkUNet", coords, feats)
    got = InferenceEngine(models[0])(torch.from_numpy(coords).cuda(), torch.from_numpy(feats).cuda())
    assert float((got.cpu().double() - want.double()).norm() / want.double().norm()) < 1e-2


def test_functional_empty_inputs():
    import lidal_b200.compat as ts
    F = ts.nn.functional
    e4 = torch.zeros((0, 4), dtype=torch.int, device="cuda")
    assert F.sphash(e4).shape == (0,)
    assert F.sphash(e4, ts.nn.utils.get_kernel_offsets(3, 1, 1, device="cuda")).shape == (27, 0)
    ref = F.sphash(torch.tensor([[1, 2, 3, 0]], dtype=torch.int, device="cuda"))
    assert F.sphashquery(torch.zeros(0, dtype=torch.int64, device="cuda"), ref).numel() == 0
    assert torch.equal(F.sphashquery(ref, torch.zeros(0, dtype=torch.int64, device="cuda")).cpu(), torch.tensor([-1]))
    assert F.spdownsample(e4, 2, 2, 1).shape == (0, 4)
    assert F.spcount(torch.zeros(0, dtype=torch.int, device="cuda"), 3).tolist() == [0, 0, 0]
    out = F.spvoxelize(torch.zeros((0, 8), device="cuda"), torch.zeros(0, dtype=torch.int, device="cuda"),
                       torch.zeros(3, dtype=torch.int, device="cuda"))
    assert out.shape == (3, 8) and float(out.abs().sum()) == 0.0


def test_scoring_frame_with_no_matches():
    """A neighbour window that is far away: zero matches -> divergence 0, entropy of the query distribution."""
    from lidal_b200 import score
    rng = np.random.default_rng(0)
    sc = score.SequenceScorer()
    xyz0 = rng.random((500, 3)) * 5
    p = rng.random((500, 19)).astype(np.float32); p /= p.sum(1, keepdims=True)
    sc.add_frame(xyz0, p, np.arange(2), [np.arange(250), np.arange(250, 500)])
    for i in range(24):
        sc.add_frame(xyz0 + 1000.0, p)
    d, e, cnt = sc.score_points(0)
    assert int(cnt.sum()) == 0 and float(d.abs().max()) == 0.0
    np.testing.assert_allclose(e.cpu().numpy(), -(p.astype(np.float64) * np.log(p.astype(np.float64))).sum(1), rtol=1e-5)
