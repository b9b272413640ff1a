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
