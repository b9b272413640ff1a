"""CPU: the oracle restatements of the steps around the network (oracle/lidal_extra.py) against the goldens produced by the
reference's OWN code in the build container (tests/golden/make_golden.py: SK_Dataset 'score' mode + collate_fn,
prepare_kdtree_sk.process_frame, segment_entropy.worker_func, ReDAL.worker_func)."""
import hashlib

import numpy as np

from lidal_b200 import synth


def sha(*arrays):
    h = hashlib.sha1()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def extra_inputs(g):
    seq = synth.make_sequence(2, "NU", seed=int(g["seq_seed"]), max_points=int(g["max_points"]))
    f = int(g["frame"])
    n, n_cls = seq.raw[f].shape[0], int(g["n_cls"])
    prob = synth.synthetic_probs(seq.xyz[f], n_cls, 9)
    outfeat = np.random.default_rng(int(g["outfeat_seed"])).normal(0, 1, (n, 96)).astype(np.float32)
    return seq, f, prob, outfeat


def test_score_batch_matches_reference_dataset_class(golden):
    import lidal_extra as ox
    g = golden["voxelizer"]
    raw = synth.raycast_scan(int(g["scan_seed"]), "NU")[:: int(g["stride"])]
    for fn in (ox.score_batch, synth.tta_batch):
        c, f, inv = fn(raw, int(g["seed"]), int(g["reps"]))
        assert c.shape[0] == int(g["n_vox"])
        assert sha(c) == str(g["coords_sha"]) and sha(f) == str(g["feats_sha"]) and sha(inv.astype(np.int64)) == str(g["inverse_sha"])
        assert np.array_equal(c[:256], g["coords_head"]) and np.array_equal(inv[:512], g["inverse_head"])


def test_register_segment_entropy_redal_match_reference(golden):
    import lidal_extra as ox
    g = golden["extra"]
    seq, f, prob, outfeat = extra_inputs(g)
    xyz = ox.register_points(seq.raw[f], g["pose"])
    assert sha(xyz) == str(g["xyz_sha"]) and np.array_equal(xyz[:64], g["xyz_head"])
    pred = np.argmax(prob, 1)
    assert ox.segment_entropy(pred, seq.sv2point[f], int(g["n_cls"])) == float(g["segment_entropy"])
    out = ox.redal_worker(prob, outfeat, g["curvature"].astype(np.float32), seq.sv_id[f], seq.sv2point[f], False)
    assert np.array_equal(out[1], g["redal_scores"]) and np.array_equal(out[2], g["redal_feats"]) and np.array_equal(out[3], g["redal_pnums"])


def test_tta_tail_matches_reference_lines(golden):
    """tail.npz: score/prob_inference.py:99-118 exec'd unmodified in the build container (gather by inverse index, softmax,
    mean over the TTA views, argmax, out_feat mean) == the oracle's tta_tail / outfeat_mean, bit for bit."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
    import lidal_extra as ox
    import lidal_scoring as orc
    from make_golden import tail_inputs
    g = golden["tail"]
    logits, feat, inverse, reps = tail_inputs()
    assert sha(logits, feat, inverse) == str(g["input_sha"])
    prob, pred = orc.tta_tail(logits, inverse, reps)
    assert prob.dtype == np.float32 and sha(prob) == str(g["prob_sha"]) and np.array_equal(prob[:64], g["prob_head"])
    assert pred.dtype == np.int64 and np.array_equal(pred, g["pred"].astype(np.int64))
    of = ox.outfeat_mean(feat, inverse, reps)
    assert of.dtype == np.float32 and sha(of) == str(g["out_feat_sha"]) and np.array_equal(of[:64], g["out_feat_head"])
