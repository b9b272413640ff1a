"""GPU: the fused inference engine against the reference-network goldens and against the module-by-module compat path."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
REL_TOL = 1e-2


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return float((a - b).norm() / b.norm())


@pytest.mark.parametrize("name,ncls", [("minkunet", 19), ("spvcnn", 16)])
def test_engine_logits_vs_reference_golden(golden, small_scan, name, ncls):
    import lidal_b200.compat as ts
    from lidal_b200.engine import InferenceEngine
    from lidal_b200.network import MinkUNet, SPVCNN, seeded_state_dict
    g = golden["nets"]
    coords, feats, _ = small_scan
    model = (MinkUNet if name == "minkunet" else SPVCNN)(ncls, ts)
    model.load_state_dict(seeded_state_dict(model.state_dict()), strict=True)
    model = model.cuda().eval()
    eng = InferenceEngine(model)
    c, f = torch.from_numpy(coords).cuda(), torch.from_numpy(feats).cuda()
    logits = eng(c, f)
    assert logits.shape == (coords.shape[0], ncls) and logits.dtype == torch.float32
    err = rel_l2(logits[:512].cpu(), torch.from_numpy(g[f"{name}_logits_head"]))
    print(f"{name} engine vs fp32 reference: rel-L2 {err:.3e}")
    assert err < REL_TOL
    with torch.no_grad():
        compat_logits = model(ts.SparseTensor(f, c))[0]
    assert rel_l2(logits.cpu(), compat_logits.cpu()) < REL_TOL
    # argmax agreement with the fp32 reference on confidently-classified rows
    ref = torch.from_numpy(g[f"{name}_logits_head"])
    top2 = ref.topk(2, 1).values
    confident = (top2[:, 0] - top2[:, 1]) > 0.05 * ref.abs().max()
    assert (logits[:512].cpu().argmax(1) == ref.argmax(1))[confident].all()


def test_engine_fp16_operands(golden, small_scan):
    import lidal_b200.compat as ts
    from lidal_b200.engine import InferenceEngine
    from lidal_b200.network import MinkUNet, seeded_state_dict
    coords, feats, _ = small_scan
    model = MinkUNet(19, ts)
    model.load_state_dict(seeded_state_dict(model.state_dict()), strict=True)
    eng = InferenceEngine(model.cuda().eval(), dtype=torch.float16)
    logits = eng(torch.from_numpy(coords).cuda(), torch.from_numpy(feats).cuda())
    err = rel_l2(logits[:512].cpu(), torch.from_numpy(golden["nets"]["minkunet_logits_head"]))
    print(f"minkunet engine fp16 operands: rel-L2 {err:.3e}")
    assert err < REL_TOL


def test_unique_and_point_queries_vs_oracle(oracle_ts, small_scan):
    import lidal_b200.compat as ts
    from lidal_b200 import _lib as L
    from lidal_b200.engine import InferenceEngine, Maps
    F, Fo = ts.nn.functional, oracle_ts.nn.functional
    coords = torch.from_numpy(small_scan[0])
    h = Fo.sphash(coords)
    keys = torch.cat([h, h[::3]])                                     # duplicates
    n = keys.numel()
    uniq = torch.empty(n, dtype=torch.int64, device="cuda")
    n_u = torch.zeros(1, dtype=torch.int, device="cuda")
    inv = torch.empty(n, dtype=torch.int, device="cuda")
    nbytes = L.lib().lb_unique_ws_bytes(n)
    ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    kd = keys.cuda()
    first = torch.empty(n, dtype=torch.int, device="cuda")
    L.check(L.lib().lb_unique_i64(L.ptr(kd), n, 60, L.ptr(uniq), L.ptr(n_u), L.ptr(inv), L.ptr(first), L.ptr(ws), nbytes, L.stream()))
    want_u, want_inv = torch.unique(keys, return_inverse=True)
    assert int(n_u.item()) == want_u.numel()
    assert torch.equal(uniq[: want_u.numel()].cpu(), want_u) and torch.equal(inv.cpu().long(), want_inv)
    _, np_first = np.unique(keys.numpy(), return_index=True)
    assert np.array_equal(first[: want_u.numel()].cpu().numpy(), np_first)
    # fused corner query == sphash(offsets) + sphashquery + calc_ti_weights
    g = torch.Generator().manual_seed(0)
    pts = torch.cat([coords[:, :3].float() + torch.rand(coords.shape[0], 3, generator=g) * 0.97, coords[:, 3:].float()], 1)
    for lvl in (0, 2):
        s = 2 ** lvl
        vox = torch.unique(torch.cat([coords[:, :3] // s * s, coords[:, 3:]], 1), dim=0).int()
        m = Maps.__new__(Maps)
        m.tables = {lvl: F._build_table(F.sphash(vox.cuda()))}
        m.n = {lvl: vox.shape[0]}
        eng = InferenceEngine.__new__(InferenceEngine)
        idx, w = InferenceEngine._corner_query(eng, pts.cuda(), m, lvl)
        off = oracle_ts.nn.utils.get_kernel_offsets(2, s, 1)
        cell = torch.cat([torch.floor(pts[:, :3] / s).int() * s, pts[:, 3:].int()], 1)
        want_idx = Fo.sphashquery(Fo.sphash(cell, off), Fo.sphash(vox))
        want_w = Fo.calc_ti_weights(pts, want_idx, scale=s)
        assert torch.equal(idx.cpu().long(), want_idx.t())
        torch.testing.assert_close(w.cpu(), want_w.t().contiguous(), rtol=1e-5, atol=1e-7)
        ci, cn = InferenceEngine._cell_query(eng, pts.cuda(), m, lvl, segments=False)
        want_ci = Fo.sphashquery(Fo.sphash(cell), Fo.sphash(vox))
        assert torch.equal(ci.cpu().long(), want_ci)
        assert torch.equal(cn.cpu(), Fo.spcount(want_ci.int(), vox.shape[0]))


def test_mask_sorted_map_and_pack8(small_scan):
    """lb_kmap_sort_by_mask: perm is a permutation ordered by neighbour mask; conv through (sorted table, out_rows)
    equals conv through the natural table; the PACK8 stem path equals the CUDA-core kernel."""
    import lidal_b200.compat as ts
    from lidal_b200 import engine
    F = ts.nn.functional
    coords = torch.from_numpy(small_scan[0]).cuda()
    n = coords.shape[0]
    nbr = F.build_kernel_map(coords, (1, 1, 1), (3, 3, 3), (1, 1, 1), (1, 1, 1)).nbr
    nbr_s, perm = engine._mask_sorted(nbr)
    assert torch.equal(torch.sort(perm.long()).values.cpu(), torch.arange(n))
    assert torch.equal(nbr_s, nbr[:, perm.long()])
    valid = nbr >= 0
    freq = valid.sum(1)
    rank = torch.argsort(torch.argsort(freq * 32 + torch.arange(27, device="cuda")))     # ascending frequency, ties by offset
    key = (valid.long() << (26 - rank).unsqueeze(1)).sum(0)                              # least frequent offset = MSB
    ms = key[perm.long()] >> 3                 # LB_MASK_KEY_BITS = 24 of the 27 key bits take part
    assert bool((ms[1:] >= ms[:-1]).all())
    # rows with equal keys keep their original order (stable)
    same = ms[1:] == ms[:-1]
    assert bool((perm[1:][same] > perm[:-1][same]).all())
    g = torch.Generator().manual_seed(0)
    x = torch.randn(n, 64, generator=g).cuda().bfloat16()
    conv = engine._Conv((torch.randn(27, 64, 96, generator=g) * 0.05).cuda(), None, relu=True)
    res = torch.randn(n, 96, generator=g).cuda().bfloat16()
    a = conv(x, nbr, n, residual=res)
    b = conv(x, (nbr_s, perm), n, residual=res)
    torch.testing.assert_close(a.float(), b.float(), rtol=1e-2, atol=1e-2)
    assert float((a.float() - b.float()).abs().mean()) < 1e-3
    # PACK8 (c_in = 4 stem) vs the SIMT kernel
    f4 = torch.randn(n, 4, generator=g).cuda()
    k4 = (torch.randn(27, 4, 32, generator=g) * 0.2).cuda()
    stem = engine._Conv(k4, None, relu=False, pack8=True)
    eng = engine.InferenceEngine.__new__(engine.InferenceEngine)
    eng.dtype = torch.bfloat16
    got = stem(eng._pad8(f4), (nbr_s, perm), n, out_dtype=torch.float32)
    want = F.conv_forward(f4.bfloat16(), F.pack_weight(k4, torch.bfloat16), nbr, n, force_simt=True)
    torch.testing.assert_close(got, want, rtol=1e-4, atol=1e-4)


def test_devoxelize16_matches_fp32_kernel():
    """The vectorised 16-bit devoxelize (engine) against the fp32 kernel on the same bf16-rounded inputs."""
    from lidal_b200 import _lib as L
    import lidal_b200.compat as ts
    g = torch.Generator().manual_seed(0)
    n, m = 20000, 3000
    for c in (32, 96, 256):
        feats = torch.randn(m, c, generator=g).cuda().bfloat16()
        idx = torch.randint(-1, m, (n, 8), generator=g).int().cuda()
        w = torch.rand(n, 8, generator=g).cuda()
        w[::3, 1:] = 0.0                                        # degenerate rows: only corner 0 contributes
        out = torch.empty(n, c, dtype=torch.bfloat16, device="cuda")
        L.check(L.lib().lb_devoxelize_fwd_ex(L.ptr(feats), L.LB_DT_BF16, c, L.ptr(idx), L.ptr(w), n, m, c, L.ptr(out),
                                             L.LB_DT_BF16, c, L.stream()))
        want = ts.nn.functional.spdevoxelize(feats.float(), idx, w)
        assert torch.equal(out, want.bfloat16())


def test_downsample_maps_equal_reference_maps_up_to_row_order(oracle_ts, small_scan):
    """lb_downsample_maps: same coarse voxel SET as spdownsample, and nbr_dn / nbr_up encode the same (child, parent,
    offset) triples as the reference's kernel map -- only the numbering of coarse rows differs."""
    from lidal_b200 import engine
    Fo = oracle_ts.nn.functional
    coords = torch.from_numpy(small_scan[0])
    m = engine.Maps(coords.cuda())
    engine_sorted = engine.SORT_MAPS
    for lvl in range(3):
        s = 2 ** lvl
        c_f = m.coords[lvl].cpu()
        nb, ns, sz, oc, res = Fo.build_kernel_map(c_f, (s, s, s), (2, 2, 2), (2, 2, 2), (1, 1, 1))
        c_c = m.coords[lvl + 1].cpu()
        key = lambda t: (t[:, 3].long() << 48) | (t[:, 0].long() << 32) | (t[:, 1].long() << 16) | t[:, 2].long()   # noqa: E731
        assert torch.equal(torch.sort(key(c_c)).values, key(oc))                 # same voxel set (oracle's is sorted)
        # map oracle coarse row -> engine coarse row
        order = torch.argsort(key(c_c))
        to_engine = order                                                         # oracle row j == engine row order[j]
        dn_s, perm = m.nbr_dn[lvl] if isinstance(m.nbr_dn[lvl], tuple) else (m.nbr_dn[lvl], None)
        dn = torch.empty_like(dn_s.cpu())
        if perm is not None:
            dn[:, perm.cpu().long()] = dn_s.cpu()
        else:                                      # strided maps are left in natural row order (engine.SORT_DN = False)
            dn = dn_s.cpu()
        assert torch.equal(dn[:, to_engine].long(), res)                          # child rows identical


def test_voxelize_ex_vec4_matches_reference_kernel():
    """float4-atomic / plain-store voxelize (engine) vs the scalar fp32 kernel: single-point and multi-point voxels."""
    from lidal_b200 import _lib as L
    import lidal_b200.compat as ts
    F = ts.nn.functional
    g = torch.Generator().manual_seed(1)
    n, m = 30000, 4000
    idx = torch.randint(-1, m, (n,), generator=g).int().cuda()
    idx[:2000] = torch.arange(2000, dtype=torch.int).cuda() + m - 2000          # plenty of single-point voxels
    idx[2000:][idx[2000:] >= m - 2000] = 5
    counts = F.spcount(idx, m)
    assert int((counts == 1).sum()) > 500 and int((counts > 4).sum()) > 500
    runs = torch.sort(torch.randint(0, m, (n,), generator=g)).values.int().cuda()     # scan-like order: long runs per voxel
    runs[::97] = -1
    for ids in (idx, runs):
        cnts = F.spcount(ids, m)
        for c in (4, 32, 256):
            feats = torch.randn(n, c, generator=g).cuda().bfloat16()
            out = torch.empty(m, c, dtype=torch.float32, device="cuda")
            L.check(L.lib().lb_voxelize_fwd_ex(L.ptr(feats), L.LB_DT_BF16, c, L.ptr(ids), L.ptr(cnts), n, m, c, L.ptr(out), L.stream()))
            want = F.spvoxelize(feats.float(), ids, cnts)
            torch.testing.assert_close(out, want, rtol=1e-5, atol=1e-5)


def test_host_pipeline_matches_direct_calls(small_scan):
    """HostPipeline (pinned H2D / D2H overlapped with compute) returns, in order, exactly what direct engine calls return."""
    import lidal_b200.compat as ts
    from lidal_b200.engine import HostPipeline, InferenceEngine
    from lidal_b200.network import MinkUNet, seeded_state_dict
    coords, feats, _ = small_scan
    model = MinkUNet(19, ts)
    model.load_state_dict(seeded_state_dict(model.state_dict()), strict=True)
    eng = InferenceEngine(model.cuda().eval())
    batches = []
    for k in range(5):
        sel = np.arange(coords.shape[0]) % (k + 2) != 0
        batches.append((torch.from_numpy(coords[sel]).pin_memory(), torch.from_numpy(feats[sel] * (1 + 0.1 * k)).pin_memory()))
    want = [eng(c.cuda(), f.cuda()).cpu().clone() for c, f in batches]
    pipe = HostPipeline(eng)
    got = []
    for c, f in batches:
        got += [o.clone() for o in pipe.submit(c, f)]
    got += [o.clone() for o in pipe.collect()]
    assert len(got) == len(want)
    for a, b in zip(got, want):
        assert a.shape == b.shape and torch.equal(a, b)


def test_accelerate_is_the_engine_behind_the_reference_call(small_scan):
    """engine.accelerate(model): the reference's own call ``logits, out_feat = model(SparseTensor(feats, coords))``
    (score/prob_inference.py:97) under no_grad returns exactly what InferenceEngine returns; a weight update is picked up."""
    import lidal_b200.compat as ts
    from lidal_b200.engine import InferenceEngine, accelerate
    from lidal_b200.network import MinkUNet, seeded_state_dict
    coords, feats, _ = small_scan
    model = MinkUNet(19, ts)          # bit-reproducible run to run (SPVCNN's point_to_voxel sums in atomic order: rounding only)
    model.load_state_dict(seeded_state_dict(model.state_dict()), strict=True)
    model = accelerate(model.cuda().eval())
    c, f = torch.from_numpy(coords).cuda(), torch.from_numpy(feats).cuda()
    want_logits, want_feat = InferenceEngine(model)(c, f, return_feat=True)
    with torch.no_grad():
        logits, feat = model(ts.SparseTensor(f, c))
        again, _ = model(ts.SparseTensor(f, c))                      # second call: cached engine
    assert logits.dtype == feat.dtype == torch.float32 and feat.shape == (coords.shape[0], 96)
    assert torch.equal(logits, want_logits) and torch.equal(feat, want_feat.float()) and torch.equal(again, logits)
    slow = model(ts.SparseTensor(f, c))[0]                            # autograd on: the module-by-module compat forward
    assert slow.requires_grad
    assert float((slow.detach() - logits).norm() / logits.norm()) < 3e-2
    with torch.no_grad():
        model.classifier[0].bias.add_(1.0)                            # in-place update bumps _version: engine rebuilt
        moved, _ = model(ts.SparseTensor(f, c))
    torch.testing.assert_close(moved, logits + 1.0, rtol=1e-5, atol=1e-4)
