"""CPU: `lidal_b200.score.cli` (the mirror of `python -m score.sv_level.LiDAL`, README.md:115) against the outputs of the
reference's own, unmodified command line run in the build container on the same synthetic Processing_files/ tree
(tests/golden/cli.npz, made by tests/golden/make_golden.py::gen_cli).  The scorer and the selector are the pinned CPU oracle
here -- what is under test is the orchestration: file discovery, flag concatenation, the per-sequence loop with its
`idx * 1000.0` centre offset, the sv_pnums / sv_centers cache, and the per-frame flag files written back."""
import glob
import os
import shutil

import numpy as np
import pytest


def _run(root, r_id, **kw):
    import lidal_scoring as orc
    from lidal_b200.score import cli
    return cli.run("SK", "SPVCNN", r_id, root=root, score_dataset=orc.score_dataset, select_regions=orc.select_regions, verbose=False, **kw)


@pytest.mark.timeout(600)
def test_cli_reproduces_the_reference_command_line(golden, tmp_path):
    from lidal_b200 import synth
    from lidal_b200.score import cli
    g = golden["cli"]
    params = dict(g["params"].tolist())
    root = str(tmp_path)
    assert synth.write_scoring_tree(root, **params) == int(g["n_regions"])
    base = os.path.join(root, "Processing_files", "SK")
    flags, paths = _run(root, 1)
    assert flags.dtype == g["flags"].dtype and np.array_equal(flags, g["flags"])
    # one int flag file per input frame, same file names, in the round's folder (LiDAL.py:140-159,327-330)
    written = sorted(glob.glob(os.path.join(base, "sv_flag", "KMeans", "SPVCNN", "LiDAL", "1r", "*", "*.npy")))
    assert written == sorted(paths) and len(written) == sum(params["lengths"].values())
    assert [len(np.load(p)) for p in written] == g["frame_sizes"].tolist()
    assert np.array_equal(np.concatenate([np.load(p) for p in written]), g["flags"])
    for seq_id in cli.SK_TRAIN_SPLIT:
        assert os.path.isdir(os.path.join(base, "sv_flag", "KMeans", "SPVCNN", "LiDAL", "1r", seq_id))
    # the region statistics cache written on the first run (LiDAL.py:220-222), centres of sequence idx offset by idx * 1000
    pn, c = (np.load(os.path.join(base, "super_voxel", "KMeans", n + ".npy")) for n in ("sv_pnums", "sv_centers"))
    assert pn.dtype == g["sv_pnums"].dtype and np.array_equal(pn, g["sv_pnums"])
    assert c.dtype == g["sv_centers"].dtype and np.array_equal(c, g["sv_centers"])
    n0 = params["lengths"]["00"] * 20
    assert c[:n0, 0].max() < 500.0 < 900.0 < c[n0:, 0].min()

    # second call of round 1: the cache is used (sv_pre, LiDAL.py:171-175) and the result is the same
    shutil.rmtree(os.path.join(base, "sv_flag", "KMeans", "SPVCNN"))
    again, _ = _run(root, 1)
    assert np.array_equal(again, g["flags"])

    # round 2 reads round 1's flags and the prob maps of the model trained in round 1 (LiDAL.py:154-155,190-191)
    for seq_id in params["lengths"]:
        shutil.copytree(os.path.join(base, "prob_map", "SPVCNN", "fr", "0r", seq_id),
                        os.path.join(base, "prob_map", "SPVCNN", "sv", "LiDAL", "1r", seq_id))
    nxt, paths2 = _run(root, 2)
    assert all(os.sep + "2r" + os.sep in p for p in paths2) and len(paths2) == len(paths)
    assert ((g["flags"] == 1) <= (nxt == 1)).all()                   # labelled regions stay labelled
    assert (nxt == 1).sum() > (g["flags"] == 1).sum()                # and the budget admits more
    with pytest.raises(AssertionError):
        _run(root, 0)                                                # LiDAL.py:150
