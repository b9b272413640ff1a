"""CPU: the oracle against the committed golden vectors (generated from the reference itself, tests/golden/make_golden.py)."""
import hashlib

import numpy as np
import torch

from lidal_b200 import synth
from lidal_b200.network import MinkUNet, SPVCNN, seeded_state_dict


def sha(*arrays):
    h = hashlib.sha1()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def test_sphash_known_answers(golden, oracle_ts):
    g = golden["hash"]
    got = oracle_ts.nn.functional.sphash(torch.from_numpy(g["coords"])).numpy()
    assert np.array_equal(got, g["hashes"])
    assert (got >= 0).all() and (got < 2 ** 60).all()


def test_kernel_offsets_order(oracle_ts):
    o3 = oracle_ts.nn.utils.get_kernel_offsets(3).tolist()
    assert o3[0] == [-1, -1, -1] and o3[1] == [0, -1, -1] and o3[13] == [0, 0, 0] and o3[26] == [1, 1, 1]
    o2 = oracle_ts.nn.utils.get_kernel_offsets(2, 4).tolist()
    assert o2 == [[0, 0, 0], [0, 0, 4], [0, 4, 0], [0, 4, 4], [4, 0, 0], [4, 0, 4], [4, 4, 0], [4, 4, 4]]


def test_scoring_oracle_matches_reference_golden(golden):
    import lidal_scoring as orc
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    from make_golden import build_scoring_inputs, scoring_case
    g = golden["scoring"]
    seq, probs = build_scoring_inputs(scoring_case())
    assert sha(*seq.xyz, *probs) == str(g["input_sha"]), "synthetic generator drifted from the committed golden inputs"
    trees = orc.build_trees(seq.xyz)
    for i in range(seq.n_frames):
        out = orc.score_frame(i, probs, seq.xyz, trees, seq.sv_id[i], seq.sv2point[i])
        for name, val in zip(("sv_id", "sv_interds", "sv_interes", "sv_pnums", "sv_centers"), out):
            assert np.array_equal(val, g[name][i]), (name, i)


def test_selection_oracle_matches_reference_golden(golden):
    import lidal_scoring as orc
    g = golden["selection"]
    out = orc.select_regions(g["sv_flags"].copy(), g["sv_interds"], g["sv_interes"], g["sv_pnums"], g["sv_centers"],
                             int(g["tight_tpn"]))
    assert np.array_equal(out, g["tight_out"])
    out = orc.select_regions(g["sv_flags"].copy(), g["loose_interds"], g["sv_interes"], g["sv_pnums"], g["sv_centers"],
                             int(g["loose_tpn"]))
    assert np.array_equal(out, g["loose_out"])
    assert set(np.unique(out)) <= {0, 1, 2}


def test_model_mirrors_match_reference_networks(golden, oracle_ts, small_scan):
    """lidal_b200.network on the oracle backend == the reference's network/*.py on the oracle (nets.npz)."""
    g = golden["nets"]
    coords, feats, _ = small_scan
    assert sha(coords, feats) == str(g["input_sha"])
    for name, cls, ncls in (("minkunet", MinkUNet, 19), ("spvcnn", SPVCNN, 16)):
        model = cls(ncls, oracle_ts)
        sd = seeded_state_dict(model.state_dict())
        assert [f"{k}:{tuple(v.shape)}" for k, v in sd.items()] == list(g[f"{name}_keys"])
        model.load_state_dict(sd, strict=True)
        model.eval()
        x = oracle_ts.SparseTensor(torch.from_numpy(feats), torch.from_numpy(coords))
        with torch.no_grad():
            logits, feat = model(x)
        assert np.array_equal(logits.numpy()[:512], g[f"{name}_logits_head"])
        assert sha(logits.numpy()) == str(g[f"{name}_logits_sha"])
        if name == "minkunet":
            for key, km in x.kmaps.items():
                tag = f"kmap_s{key[0][0]}_k{key[1][0]}_st{key[2][0]}"
                assert np.array_equal(km[1].numpy(), g[tag + "_nbsizes"])
                assert sha(km[0].numpy().astype(np.int32)) == str(g[tag + "_sha"])


def test_kernel_map_invariants(oracle_ts, small_scan):
    F = oracle_ts.nn.functional
    coords = torch.from_numpy(small_scan[0])
    nbmaps, nbsizes, sizes, out_coords, results = F.build_kernel_map(coords, (1, 1, 1), (3, 3, 3), (1, 1, 1), (1, 1, 1))
    n = coords.shape[0]
    assert sizes == (n, n) and int(nbsizes[13]) == n            # centre offset: every voxel is its own neighbour
    assert torch.equal(results[13], torch.arange(n))
    assert torch.equal(results[0] >= 0, torch.isin(torch.arange(n), results[26][results[26] >= 0]) | (results[0] >= 0))
    # symmetry: o has neighbour i at offset k  <=>  i has neighbour o at offset 26-k
    k = 5
    o = torch.nonzero(results[k] >= 0).squeeze(1)
    assert torch.equal(results[26 - k][results[k][o]], o)
    _, _, sizes2, oc2, res2 = F.build_kernel_map(coords, (1, 1, 1), (2, 2, 2), (2, 2, 2), (1, 1, 1))
    assert int((res2 >= 0).sum()) == n                              # every fine voxel has exactly one parent
    packed = (oc2[:, 3].long() << 48) | (oc2[:, 0].long() << 32) | (oc2[:, 1].long() << 16) | oc2[:, 2].long()
    assert bool((packed[1:] > packed[:-1]).all())                   # sorted by (b, x, y, z), unique


def _dense(feats, coords, size, batches):
    """Scatter a sparse tensor's rows into a dense [B, C, X, Y, Z] grid (zeros where no voxel is active)."""
    vol = torch.zeros(batches, feats.shape[1], size, size, size, dtype=feats.dtype)
    c = coords.long()
    vol[c[:, 3], :, c[:, 0], c[:, 1], c[:, 2]] = feats
    return vol


def test_conv_oracle_equals_dense_conv3d(oracle_ts):
    """Independent pin of the conv arithmetic: the oracle's gather -> mm -> scatter-add loop (submanifold k3, strided k2s2 and
    its transposed use, network/utils.py:110-133) equals torch's DENSE conv3d / conv_transpose3d on the densified grid, read
    back at the active sites, with weight[k] placed at kernel tap get_kernel_offsets()[k].  (What this cannot pin is the
    third-party offset order itself -- that stays 'TS-recalled', oracle/README.md.)"""
    ts = oracle_ts
    F = ts.nn.functional
    rng = np.random.default_rng(3)
    size, batches, cin, cout = 12, 2, 5, 7
    c = np.unique(np.c_[rng.integers(1, size - 1, (500, 3)), rng.integers(0, batches, 500)], axis=0).astype(np.int32)
    coords = torch.from_numpy(c)
    feats = torch.from_numpy(rng.normal(size=(c.shape[0], cin)).astype(np.float32))
    x = ts.SparseTensor(feats, coords)
    grid = _dense(feats, coords, size, batches)

    def taps(w, ks, transposed=False):
        """[K, Cin, Cout] in offset order -> dense [Cout, Cin, kx, ky, kz] ([Cin, Cout, ...] for conv_transpose3d)."""
        offs = ts.nn.utils.get_kernel_offsets(ks).long()
        offs = offs - offs.min(0).values
        dense = torch.zeros((w.shape[1], w.shape[2]) if transposed else (w.shape[2], w.shape[1]), dtype=w.dtype)
        dense = dense[..., None, None, None].repeat(1, 1, ks, ks, ks)
        for k, (dx, dy, dz) in enumerate(offs.tolist()):
            dense[:, :, dx, dy, dz] = w[k] if transposed else w[k].t()
        return dense

    # submanifold 3x3x3, stride 1: outputs exist only at the active input sites
    w3 = torch.from_numpy(rng.normal(size=(27, cin, cout)).astype(np.float32))
    y = F.conv3d(x, w3, 3)
    want = torch.nn.functional.conv3d(grid, taps(w3, 3), padding=1)
    cl = coords.long()
    torch.testing.assert_close(y.feats, want[cl[:, 3], :, cl[:, 0], cl[:, 1], cl[:, 2]], rtol=1e-4, atol=1e-5)
    assert torch.equal(y.coords, coords)

    # 2x2x2 stride 2: outputs at the occupied coarse cells, sorted by (b, x, y, z)
    w2 = torch.from_numpy(rng.normal(size=(8, cin, cout)).astype(np.float32))
    z = F.conv3d(x, w2, 2, stride=2)
    want = torch.nn.functional.conv3d(grid, taps(w2, 2), stride=2)
    zc = z.coords.long()
    assert z.stride == (2, 2, 2) and bool((zc[:, :3] % 2 == 0).all())
    torch.testing.assert_close(z.feats, want[zc[:, 3], :, zc[:, 0] // 2, zc[:, 1] // 2, zc[:, 2] // 2], rtol=1e-4, atol=1e-5)
    occupied = torch.unique(torch.cat([torch.div(cl[:, :3], 2, rounding_mode="floor") * 2, cl[:, 3:]], 1), dim=0)
    assert zc.shape[0] == occupied.shape[0]

    # transposed 2x2x2 stride 2 back to the fine level: the cached map with its roles swapped == conv_transpose3d
    wt = torch.from_numpy(rng.normal(size=(8, cout, cin)).astype(np.float32))
    u = F.conv3d(z, wt, 2, stride=2, transposed=True)
    coarse = torch.zeros(batches, cout, size // 2, size // 2, size // 2)
    coarse[zc[:, 3], :, zc[:, 0] // 2, zc[:, 1] // 2, zc[:, 2] // 2] = z.feats
    want = torch.nn.functional.conv_transpose3d(coarse, taps(wt, 2, transposed=True), stride=2)
    assert torch.equal(u.coords, coords) and u.stride == (1, 1, 1)
    torch.testing.assert_close(u.feats, want[cl[:, 3], :, cl[:, 0], cl[:, 1], cl[:, 2]], rtol=1e-4, atol=1e-5)


def test_devoxelize_oracle_equals_dense_trilinear_sampling(oracle_ts):
    """Independent pin of the point branch's arithmetic: on a fully occupied grid (no missing corners, so no renormalisation)
    the oracle's voxel_to_point chain of network/utils.py:66-83 -- 8-corner hash query, calc_ti_weights, spdevoxelize -- is
    torch's dense trilinear ``grid_sample`` (align_corners=True), at tensor strides 1 and 2; and point_to_voxel's
    scatter-mean (network/utils.py:38-61) is a per-cell mean."""
    ts = oracle_ts
    F = ts.nn.functional
    rng = np.random.default_rng(8)
    cells, ch = 6, 3
    for stride in (1, 2):
        g = np.stack(np.meshgrid(*[np.arange(cells)] * 3, indexing="ij"), -1).reshape(-1, 3) * stride
        vc = torch.from_numpy(np.c_[g, np.zeros(len(g), int)].astype(np.int32))
        vf = torch.from_numpy(rng.normal(size=(len(g), ch)).astype(np.float32))
        pts = torch.from_numpy(rng.uniform(0.0, (cells - 1) * stride - 1e-3, (200, 3)).astype(np.float32))
        zc = torch.cat([pts, torch.zeros(200, 1)], 1)
        off = ts.nn.utils.get_kernel_offsets(2, (stride,) * 3, 1)
        old_hash = F.sphash(torch.cat([torch.floor(zc[:, :3] / stride).int() * stride, zc[:, -1].int().view(-1, 1)], 1), off)
        idx_query = F.sphashquery(old_hash, F.sphash(vc))
        assert bool((idx_query >= 0).all())
        weights = F.calc_ti_weights(zc, idx_query, scale=stride).transpose(0, 1).contiguous()
        got = F.spdevoxelize(vf, idx_query.transpose(0, 1).contiguous(), weights)
        vol = vf.t().reshape(1, ch, cells, cells, cells)                         # [B, C, X, Y, Z]
        norm = pts / stride / (cells - 1) * 2 - 1
        want = torch.nn.functional.grid_sample(vol, norm[:, [2, 1, 0]].view(1, 1, 1, -1, 3), mode="bilinear", align_corners=True)
        torch.testing.assert_close(got, want.view(ch, -1).t(), rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(weights.sum(1), torch.ones(200), rtol=1e-5, atol=1e-5)
        # scatter-mean: points -> the cell that contains them
        cell = torch.floor(pts / stride).long()
        lin = (cell[:, 0] * cells + cell[:, 1]) * cells + cell[:, 2]
        pf = torch.from_numpy(rng.normal(size=(200, ch)).astype(np.float32))
        pc_hash = F.sphash(torch.cat([cell.int() * stride, torch.zeros(200, 1, dtype=torch.int)], 1))
        idx = F.sphashquery(pc_hash, F.sphash(vc))
        assert torch.equal(idx, lin)                                             # grid rows were laid out x-major
        counts = F.spcount(idx.int(), len(g))
        mean = F.spvoxelize(pf, idx, counts)
        want = torch.zeros(len(g), ch).index_add_(0, lin, pf) / torch.bincount(lin, minlength=len(g)).clamp_min(1)[:, None]
        torch.testing.assert_close(mean, want, rtol=1e-5, atol=1e-6)
        assert torch.equal(counts.long(), torch.bincount(lin, minlength=len(g)))
