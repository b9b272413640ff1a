"""GPU: SPVCNN point<->voxel kernels vs the oracle (fp32, tolerance 1e-5 relative; indices bit-exact)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ts():
    import lidal_b200.compat as ts
    return ts


def test_voxelize_devoxelize_roundtrip_vs_oracle(ts, oracle_ts, small_scan):
    from lidal_b200.network.point_voxel import PointVoxel
    coords, feats, _ = small_scan
    g = torch.Generator().manual_seed(0)
    f32 = torch.randn(coords.shape[0], 32, generator=g)
    jitter = torch.rand(coords.shape[0], 3, generator=g) * 0.98
    outs = {}
    for name, be, dev in (("o", oracle_ts, "cpu"), ("g", ts, "cuda")):
        pv = PointVoxel(be)
        pc = torch.cat([torch.from_numpy(coords[:, :3]).float() + jitter, torch.from_numpy(coords[:, 3:]).float()], 1)
        z = be.PointTensor(f32.to(dev), pc.to(dev))
        # resolution 1.0: (C * r) / r is exact on both devices.  With r = 0.05 CUDA torch evaluates the caller's
        # `x / 0.05` as `x * (1 / 0.05)` (1 ulp off CPU torch at |x| ~ 8192) -- outside the boundary under test.
        x0 = pv.initial_voxelize(z, 1.0, 1.0)
        z0 = pv.voxel_to_point(x0, z)
        x1 = pv.point_to_voxel(x0, z0)
        # a strided level: coarse voxels at stride 4
        cs = torch.unique(torch.cat([x0.C[:, :3] // 4 * 4, x0.C[:, 3:]], 1), dim=0)
        xs = be.SparseTensor(torch.randn(cs.shape[0], 16, generator=torch.Generator().manual_seed(5)).to(dev), cs.int(), 4)
        z4 = pv.voxel_to_point(xs, z0)
        outs[name] = dict(C=x0.C.cpu(), F=x0.F.cpu(), z0=z0.F.cpu(), x1=x1.F.cpu(), z4=z4.F.cpu(),
                          iq=z0.idx_query[(1, 1, 1)].cpu(), w=z0.weights[(1, 1, 1)].cpu(),
                          iq4=z4.idx_query[(4, 4, 4)].cpu(), w4=z4.weights[(4, 4, 4)].cpu())
    o, g_ = outs["o"], outs["g"]
    assert torch.equal(o["C"], g_["C"]) and torch.equal(o["iq"], g_["iq"]) and torch.equal(o["iq4"], g_["iq4"])
    for k in ("F", "z0", "x1", "z4", "w", "w4"):
        torch.testing.assert_close(g_[k], o[k], rtol=1e-5, atol=2e-6, msg=lambda m: f"{k}: {m}")


def test_point_voxel_backward_vs_oracle(ts, oracle_ts):
    g = torch.Generator().manual_seed(2)
    n, m, c = 5000, 700, 24
    idx = torch.randint(-1, m, (n,), generator=g)
    feats = torch.randn(n, c, generator=g)
    outs = []
    for be, dev in ((oracle_ts, "cpu"), (ts, "cuda")):
        F = be.nn.functional
        counts = F.spcount(idx.int().to(dev), m)
        f = feats.clone().to(dev).requires_grad_(True)
        v = F.spvoxelize(f, idx.to(dev), counts)
        i8 = torch.randint(-1, m, (n, 8), generator=torch.Generator().manual_seed(4)).to(dev)
        w8 = torch.rand(n, 8, generator=torch.Generator().manual_seed(6)).to(dev)
        p = F.spdevoxelize(v, i8, w8)
        p.square().sum().backward()
        outs.append((v.detach().cpu(), p.detach().cpu(), f.grad.cpu()))
    for a, b in zip(outs[1], outs[0]):
        torch.testing.assert_close(a, b, rtol=2e-4, atol=1e-4)


def test_tta_tail_vs_oracle():
    import lidal_scoring as orc
    from lidal_b200 import score
    rng = np.random.default_rng(0)
    n_pts, n_vox_per, reps, ncls = 20000, 15000, 8, 19
    logits = (rng.normal(size=(reps * n_vox_per, ncls)) * 3).astype(np.float32)
    inv = np.concatenate([rng.integers(0, n_vox_per, n_pts) + v * n_vox_per for v in range(reps)]).astype(np.int64)
    want_p, want_y = orc.tta_tail(logits, inv, reps)
    got_p, got_y = score.tta_tail(torch.from_numpy(logits).cuda(), torch.from_numpy(inv).cuda(), reps)
    got_p = got_p.cpu().numpy()
    np.testing.assert_allclose(got_p, want_p, rtol=1e-5, atol=1e-7)
    np.testing.assert_allclose(got_p.sum(1), 1.0, atol=1e-5)
    agree = (got_y.cpu().numpy() == want_y)
    # argmax may only differ where the top-2 probabilities are within float rounding of each other
    top2 = np.sort(want_p, 1)[:, -2:]
    assert (agree | (top2[:, 1] - top2[:, 0] < 1e-6)).all() and agree.mean() > 0.9999


def test_gpu_voxelizer_matches_reference_transform():
    """F1: device TTA views == dataset/sk_dataset.py:143-169 (restated in lidal_b200.synth.score_transform) with the same
    RandomState: voxel coordinates, unique order, first-point rule and inverse indices bit-exact; features to 1 f32 ulp."""
    from lidal_b200 import synth, voxelizer
    raw = synth.raycast_scan(11, "NU")
    want_c, want_f, want_inv = synth.tta_batch(raw, seed=5, inf_reps=3)
    got_c, got_f, got_inv = voxelizer.tta_batch_gpu(torch.from_numpy(raw).cuda(), seed=5, inf_reps=3)
    assert np.array_equal(got_c.cpu().numpy(), want_c)
    assert np.array_equal(got_inv.cpu().numpy(), want_inv)
    np.testing.assert_allclose(got_f.cpu().numpy(), want_f, rtol=2e-7, atol=1e-6)


def test_segmented_voxelize_matches_scatter_mean():
    """lb_segment_order + lb_voxelize_segments (engine: atomic-free point_to_voxel) == the scatter-mean of F.spvoxelize on the
    same 16-bit features: unmatched points (-1) are skipped, empty voxels stay zero, the order list is a partition."""
    from lidal_b200 import _lib as L
    g = torch.Generator().manual_seed(3)
    n, m = 50_000, 4_000
    for c in (32, 128, 256):
        idx = torch.randint(-1, m, (n,), generator=g).int()
        idx[idx == 7] = -1                                       # voxel 7 stays empty
        feats = torch.randn(n, c + 8, generator=g).bfloat16()[:, :c]           # strided rows
        idx_d, feats_d = idx.cuda(), feats.cuda()
        counts = torch.empty(m, dtype=torch.int, device="cuda")
        L.check(L.lib().lb_count(L.ptr(idx_d), n, L.ptr(counts), m, L.stream()))
        seg = torch.empty(m + 1, dtype=torch.int, device="cuda")
        order = torch.full((n,), -7, dtype=torch.int, device="cuda")
        nb = L.lib().lb_segment_order_ws_bytes(m)
        ws = torch.empty(nb, dtype=torch.uint8, device="cuda")
        L.check(L.lib().lb_segment_order(L.ptr(idx_d), n, L.ptr(counts), m, L.ptr(seg), L.ptr(order), L.ptr(ws), nb, L.stream()))
        seg_h, order_h = seg.cpu().numpy(), order.cpu().numpy()
        total = int(seg_h[-1])
        assert total == int((idx >= 0).sum()) and np.array_equal(np.diff(seg_h), counts.cpu().numpy())
        assert np.array_equal(np.sort(order_h[:total]), np.nonzero((idx >= 0).numpy())[0])
        assert np.array_equal(idx.numpy()[order_h[:total]], np.repeat(np.arange(m), np.diff(seg_h)))
        out = torch.empty((m, c), dtype=torch.bfloat16, device="cuda")
        L.check(L.lib().lb_voxelize_segments(L.ptr(feats_d), L.LB_DT_BF16, feats_d.stride(0), L.ptr(order), L.ptr(seg), m, c, L.ptr(out),
                                             c, L.stream()))
        want = torch.zeros(m, c, dtype=torch.float64)
        sel = idx >= 0
        want.index_add_(0, idx[sel].long(), feats[sel].double())
        want = want / counts.cpu().double().clamp_min(1)[:, None]
        torch.testing.assert_close(out.cpu().double(), want, rtol=1e-2, atol=1e-2)
        assert float(out[7].abs().max()) == 0.0
