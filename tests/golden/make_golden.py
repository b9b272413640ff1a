"""Generate tests/golden/*.npz in the BUILD container (needs /root/reference; never runs on the GPU box).

What it pins:
  scoring.npz    the reference's own ``score.sv_level.LiDAL.worker_func`` executed on synthetic
                 prob / KD-tree / supervoxel files  == oracle.lidal_scoring.score_frame (exact).
  selection.npz  the reference's own selection source (LiDAL.py:230-325, exec'd unmodified with a
                 small ``train_point_num`` so the budget binds) == oracle.lidal_scoring.select_regions.
  nets.npz       the reference's unmodified network/minkunet.py + network/spvcnn.py running on the
                 oracle torchsparse restatement: logits slices + checksums, kernel-map sizes/checksums.
  hash.npz       sphash known answers from an independent pure-Python-int FNV-1a.

  voxelizer.npz  the reference's own ``dataset.sk_dataset.SK_Dataset`` ('score' mode, ``__getitem__`` x 8 + ``collate_fn``) on a
                 synthetic .bin file under ``np.random.seed`` == oracle.lidal_extra.score_batch == lidal_b200.synth.tta_batch.
  extra.npz      the reference's own ``dataset.prepare_kdtree_sk.process_frame`` (registered coordinates read back from the
                 KD-tree pickle), ``score.frame_level.segment_entropy.worker_func`` and ``score.sv_level.ReDAL.worker_func``
                 on synthetic files == oracle.lidal_extra.{register_points, segment_entropy, redal_worker} (exact).

  cli.npz        the reference's own command line ``python -m score.sv_level.LiDAL --dataset_name SK --model_name SPVCNN --r_id 1``
                 run unmodified in a synthetic Processing_files/ tree == lidal_b200.score.cli.run on the oracle: every
                 sv_flag/*.npy file and the sv_pnums / sv_centers cache byte-identical.

Run:  python tests/golden/make_golden.py
"""
import hashlib
import io
import os
import pickle
import sys
import tempfile
from contextlib import redirect_stdout

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
REF = "/root/reference"
sys.path[:0] = [ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "oracle", "_stubs"), REF]
OUT = os.path.dirname(os.path.abspath(__file__))

from lidal_b200 import synth  # noqa: E402
import lidal_scoring as orc  # noqa: E402


def sha(*arrays):
    h = hashlib.sha1()
    for a in arrays:
        h.update(np.ascontiguousarray(a).tobytes())
    return h.hexdigest()


def scoring_case():
    """26 frames (minimum for the +-12 reflection rule is 25), 1500 pts, 19 classes."""
    return dict(n_frames=26, kind="NU", seed=11, max_points=1500, n_cls=19, step=0.15)


def build_scoring_inputs(case):
    seq = synth.make_sequence(case["n_frames"], case["kind"], case["seed"], max_points=case["max_points"],
                              step=case["step"])
    probs = [synth.synthetic_probs(seq.xyz[i], case["n_cls"], 500 + i) for i in range(seq.n_frames)]
    return seq, probs


def gen_scoring():
    from sklearn.neighbors import KDTree
    import score.sv_level.LiDAL as ref            # the reference itself (nuscenes stub on sys.path)
    case = scoring_case()
    seq, probs = build_scoring_inputs(case)
    with tempfile.TemporaryDirectory() as d:
        pf, kf, sf = [], [], []
        for i in range(seq.n_frames):
            pf.append(f"{d}/p{i:06d}.npy"); np.save(pf[-1], probs[i])
            kf.append(f"{d}/k{i:06d}.pickle")
            with open(kf[-1], "wb") as f:
                pickle.dump(KDTree(seq.xyz[i]), f)
            sf.append(f"{d}/s{i:06d}.pickle")
            with open(sf[-1], "wb") as f:
                pickle.dump((seq.sv_id[i], seq.sv2point[i]), f)
        ref.init_worker(False, 24, 0.1, "00", pf, kf, sf)
        with redirect_stdout(io.StringIO()):
            ref_out = [ref.worker_func(i) for i in range(seq.n_frames)]
    trees = orc.build_trees(seq.xyz)
    for i in range(seq.n_frames):
        mine = orc.score_frame(i, probs, seq.xyz, trees, seq.sv_id[i], seq.sv2point[i])
        for a, b in zip(ref_out[i], mine):
            assert a.dtype == b.dtype and np.array_equal(a, b), f"oracle != reference at frame {i}"
    # per-point values for finer-grained GPU parity
    pts = []
    for i in (0, 13, 25):
        nids = orc.neighbour_ids(i, seq.n_frames)
        d_, e_, c_ = orc.score_points(probs[i], seq.xyz[i], [probs[n] for n in nids], [trees[n] for n in nids])
        pts.append((d_, e_, c_))
    np.savez_compressed(
        f"{OUT}/scoring.npz", case=np.array(list(case.items()), dtype=object),
        input_sha=sha(*seq.xyz, *probs),
        sv_id=np.stack([r[0] for r in ref_out]), sv_interds=np.stack([r[1] for r in ref_out]),
        sv_interes=np.stack([r[2] for r in ref_out]), sv_pnums=np.stack([r[3] for r in ref_out]),
        sv_centers=np.stack([r[4] for r in ref_out]),
        pt_frames=np.array([0, 13, 25]),
        **{f"pt_d{j}": p[0] for j, p in enumerate(pts)}, **{f"pt_e{j}": p[1] for j, p in enumerate(pts)},
        **{f"pt_c{j}": p[2] for j, p in enumerate(pts)}, allow_pickle=True)
    print("scoring.npz: oracle == reference worker_func on", seq.n_frames, "frames")
    return ref_out


def selection_inputs(seed=5, n=1200):
    rng = np.random.default_rng(seed)
    sv_flags = (rng.random(n) < 0.02).astype(np.float64)              # np.append result dtype
    sv_flags[rng.random(n) < 0.05] = 2.0                               # stale pseudo labels from last round
    sv_interds = rng.gamma(2.0, 0.05, n).astype(np.float32)
    sv_interds[rng.random(n) < 0.03] = 0.0
    sv_interes = rng.random(n).astype(np.float32)
    sv_pnums = rng.integers(60, 80, n).astype(int)
    t = np.linspace(0, 300, n)
    sv_centers = np.stack([t + rng.normal(0, 3, n), rng.normal(0, 6, n), rng.normal(0, 0.5, n)], 1).astype(np.float32)
    return sv_flags, sv_interds, sv_interes, sv_pnums, sv_centers


def gen_selection():
    src = open(f"{REF}/score/sv_level/LiDAL.py").read().split("\n")
    block = "\n".join(line[4:] if line.startswith("    ") else line for line in src[224:325])   # lines 225..325
    cases = {}
    for name, tpn in (("tight", 700_000), ("loose", 100_000_000)):
        sv_flags, d, e, pn, c = selection_inputs()
        if name == "loose":
            # np.argsort's order among exact ties (D == 0) is an accident of its quicksort; the loose case walks every
            # candidate, so give it distinct keys.  The tight case keeps the zeros (SL pass skips them, LiDAL.py:298).
            z = np.where(d == 0)[0]
            d = d.copy(); d[z] = (1e-6 * (1 + np.arange(len(z)))).astype(np.float32)
        ns = dict(np=np, sv_flags=sv_flags.copy(), sv_interds=d, sv_interes=e, sv_pnums=pn, sv_centers=c,
                  train_point_num=tpn)
        with redirect_stdout(io.StringIO()):
            exec(compile(block, "LiDAL.py[225:325]", "exec"), ns)
        mine = orc.select_regions(sv_flags.copy(), d, e, pn, c, tpn)
        assert np.array_equal(ns["sv_flags"], mine), name
        cases[name] = (tpn, ns["sv_flags"], d)
        print(f"selection[{name}]: oracle == reference block; flags1={int((mine == 1).sum())} flags2={int((mine == 2).sum())}")
    sv_flags, d, e, pn, c = selection_inputs()
    np.savez_compressed(f"{OUT}/selection.npz", sv_flags=sv_flags, sv_interds=d, sv_interes=e, sv_pnums=pn,
                        sv_centers=c, tight_tpn=cases["tight"][0], tight_out=cases["tight"][1],
                        loose_tpn=cases["loose"][0], loose_out=cases["loose"][1], loose_interds=cases["loose"][2])


def fnv_python(c4):
    h = 14695981039346656037
    for v in c4:
        h ^= (int(v) & 0xFFFFFFFF)
        h = (h * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return (h >> 60) ^ (h & 0x0FFFFFFFFFFFFFFF)


def gen_hash():
    import torch
    import torchsparse.nn.functional as F
    rng = np.random.default_rng(3)
    coords = np.concatenate([rng.integers(0, 8192, (12, 3)), rng.integers(0, 8, (12, 1))], 1).astype(np.int32)
    coords = np.concatenate([coords, [[0, 0, 0, 0], [8191, 8191, 8191, 7], [-1, -2, -3, 0], [1, 0, 0, 0]]], 0).astype(np.int32)
    want = np.array([fnv_python(c) for c in coords], dtype=np.int64)
    got = F.sphash(torch.from_numpy(coords)).numpy()
    assert np.array_equal(want, got)
    np.savez_compressed(f"{OUT}/hash.npz", coords=coords, hashes=want)
    print("hash.npz: oracle sphash == pure-python FNV-1a on 16 coords")


def gen_nets():
    import torch
    import torchsparse
    from torchsparse import SparseTensor
    from network.minkunet import MinkUNet          # reference, unmodified
    from network.spvcnn import SPVCNN
    from lidal_b200.network import seeded_state_dict
    raw = synth.raycast_scan(42, "NU")
    rs = np.random.RandomState(9)
    coords, feats, inv = synth.collate_views([synth.score_transform(raw[::4], rs), synth.score_transform(raw[1::4], rs)])
    out = dict(input_sha=sha(coords, feats), n_vox=coords.shape[0])
    for name, cls, ncls in (("minkunet", MinkUNet, 19), ("spvcnn", SPVCNN, 16)):
        model = cls(ncls)
        sd = seeded_state_dict(model.state_dict(), seed=7122)
        model.load_state_dict(sd, strict=True)
        model.eval()
        x = SparseTensor(torch.from_numpy(feats), torch.from_numpy(coords))
        with torch.no_grad():
            logits, feat = model(x)
        logits, feat = logits.numpy(), feat.numpy()
        out[f"{name}_keys"] = np.array([f"{k}:{tuple(v.shape)}" for k, v in sd.items()])
        out[f"{name}_logits_head"] = logits[:512]
        out[f"{name}_logits_sha"] = sha(logits)
        out[f"{name}_logits_absmean"] = np.abs(logits).mean()
        out[f"{name}_feat_head"] = feat[:64]
        if name == "minkunet":
            for key, km in x.kmaps.items():
                tag = f"kmap_s{key[0][0]}_k{key[1][0]}_st{key[2][0]}"
                out[tag + "_nbsizes"] = km[1].numpy()
                out[tag + "_sha"] = sha(km[0].numpy().astype(np.int32))
                out[tag + "_sizes"] = np.array(km[2])
            for s, c in x.cmaps.items():
                out[f"cmap_s{s[0]}_sha"] = sha(c.numpy().astype(np.int32))
        print(f"nets.npz: {name} logits {logits.shape} |mean|={np.abs(logits).mean():.4f}")
    np.savez_compressed(f"{OUT}/nets.npz", **out)


def gen_voxelizer():
    import lidal_extra as ox
    from dataset.sk_dataset import SK_Dataset            # the reference's dataset class, unmodified
    raw = synth.raycast_scan(21, "NU")[::3]
    seed, reps = 3, 8
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(f"{d}/Processing_files/SK")
        os.makedirs(f"{d}/sequences/00/velodyne")
        path = f"{d}/sequences/00/velodyne/000000.bin"
        raw.tofile(path)
        os.chdir(d)
        try:
            ds = SK_Dataset("score", [path] * reps)
            np.random.seed(seed)
            batch = ds.collate_fn([ds[i] for i in range(reps)])
        finally:
            os.chdir(cwd)
    coords, feats, inv = batch["coords_v_b"].numpy(), batch["feats_v_b"].numpy(), batch["inverse_indices_b"].numpy()
    for name, fn in (("oracle.lidal_extra.score_batch", ox.score_batch), ("lidal_b200.synth.tta_batch", synth.tta_batch)):
        c2, f2, i2 = fn(raw, seed, reps)
        assert c2.dtype == coords.dtype and np.array_equal(c2, coords), name
        assert f2.dtype == feats.dtype and np.array_equal(f2, feats), name
        assert np.array_equal(i2, inv), name
    np.savez_compressed(f"{OUT}/voxelizer.npz", scan_seed=21, stride=3, seed=seed, reps=reps, n_vox=coords.shape[0],
                        coords_sha=sha(coords), feats_sha=sha(feats), inverse_sha=sha(inv.astype(np.int64)),
                        coords_head=coords[:256], feats_head=feats[:256], inverse_head=inv[:512])
    print(f"voxelizer.npz: reference SK_Dataset('score') + collate_fn == oracle == synth on {raw.shape[0]} points, {coords.shape[0]} voxels")


def gen_extra():
    import lidal_extra as ox
    import dataset.prepare_kdtree_sk as ref_kd
    import score.frame_level.segment_entropy as ref_se
    import score.sv_level.ReDAL as ref_redal
    rng = np.random.default_rng(17)
    seq = synth.make_sequence(2, "NU", seed=4, max_points=6000)
    raw, pose = seq.raw[1], seq.poses[1]
    sv_id, sv2point = seq.sv_id[1], seq.sv2point[1]
    n, n_cls = raw.shape[0], 16
    prob = synth.synthetic_probs(seq.xyz[1], n_cls, 9)
    pred = np.argmax(prob, 1)
    outfeat = rng.normal(0, 1, (n, 96)).astype(np.float32)
    curvature = rng.random(n)                                  # float64 on disk; ReDAL.py:57 casts to float32
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(f"{d}/Processing_files/SK/kdtree/00")
        os.makedirs(f"{d}/sequences/00/velodyne")
        path = f"{d}/sequences/00/velodyne/000001.bin"
        raw.tofile(path)
        os.chdir(d)
        try:
            with redirect_stdout(io.StringIO()):
                ref_kd.process_frame(0, [path], [pose])
            with open("Processing_files/SK/kdtree/00/000001.pickle", "rb") as f:
                xyz_ref = np.asarray(pickle.load(f).data)
        finally:
            os.chdir(cwd)
        np.save(f"{d}/prob.npy", prob); np.save(f"{d}/pred.npy", pred); np.save(f"{d}/feat.npy", outfeat); np.save(f"{d}/curv.npy", curvature)
        with open(f"{d}/sv.pickle", "wb") as f:
            pickle.dump((sv_id, sv2point), f)
        ref_se.init_worker(n_cls, "00", [f"{d}/pred.npy"], [f"{d}/sv.pickle"])
        with redirect_stdout(io.StringIO()):
            sege_ref = ref_se.worker_func(0)
        ref_redal.init_worker(False, "00", [f"{d}/prob.npy"], [f"{d}/feat.npy"], [f"{d}/curv.npy"], [f"{d}/sv.pickle"])
        with redirect_stdout(io.StringIO()):
            redal_ref = ref_redal.worker_func(0)
    xyz = ox.register_points(raw, pose)
    assert xyz.dtype == xyz_ref.dtype and np.array_equal(xyz, xyz_ref), "register_points"
    assert np.array_equal(xyz, seq.xyz[1]), "synth.make_sequence registration"
    sege = ox.segment_entropy(pred, sv2point, n_cls)
    assert sege == sege_ref, "segment_entropy"
    mine = ox.redal_worker(prob, outfeat, curvature.astype(np.float32), sv_id, sv2point, False)
    for a, b in zip(redal_ref, mine):
        assert a.dtype == b.dtype and np.array_equal(a, b), "redal_worker"
    np.savez_compressed(f"{OUT}/extra.npz", seq_seed=4, frame=1, max_points=6000, n_cls=n_cls, pose=pose, xyz_sha=sha(xyz_ref),
                        xyz_head=xyz_ref[:64], segment_entropy=sege_ref, outfeat_seed=17, curvature=curvature,
                        redal_scores=redal_ref[1], redal_feats=redal_ref[2], redal_pnums=redal_ref[3])
    print(f"extra.npz: reference process_frame / segment_entropy / ReDAL worker == oracle on {n} points; sege={sege_ref:.6f}")


def tail_inputs(seed=77, reps=8, n_pts=700, n_cls=19):
    """Ragged TTA views: per view its own voxel count, every point of the scan mapped into every view (collate_fn :216-218,230)."""
    rng = np.random.default_rng(seed)
    n_vox = [int(v) for v in rng.integers(400, 600, reps)]
    logits = (rng.normal(size=(sum(n_vox), n_cls)) * 3).astype(np.float32)
    feat = rng.normal(size=(sum(n_vox), 96)).astype(np.float32)
    inverse = np.concatenate([rng.integers(0, nv, n_pts) + off for nv, off in zip(n_vox, np.cumsum([0] + n_vox[:-1]))]).astype(np.int64)
    return logits, feat, inverse, reps


def gen_tail():
    """The reference's own inference tail -- score/prob_inference.py:99-118 exec'd unmodified (the model call above it needs
    CUDA, these lines do not): gather by inverse index, softmax, mean over the TTA views, argmax, out_feat mean ==
    oracle.lidal_scoring.tta_tail / oracle.lidal_extra.outfeat_mean (exact)."""
    import types
    import torch
    import lidal_extra as ox
    src = open(f"{REF}/score/prob_inference.py").read().split("\n")
    block = "\n".join(line[12:] if line.startswith(" " * 12) else line.strip() for line in src[98:118])      # lines 99..118
    assert block.lstrip().startswith("# Project to original points") and "out_feat = np.mean(out_feat, axis=0)" in block
    logits, feat, inverse, reps = tail_inputs()
    ns = dict(torch=torch, np=np, logits_v_b=torch.from_numpy(logits), out_feat_v_b=torch.from_numpy(feat),
              batch={"inverse_indices_b": torch.from_numpy(inverse)}, args=types.SimpleNamespace(r_id=0, metric_name="LiDAL", inf_reps=reps))
    exec(compile(block, "prob_inference.py[99:118]", "exec"), ns)
    prob, pred = orc.tta_tail(logits, inverse, reps)
    of = ox.outfeat_mean(feat, inverse, reps)
    for name, a, b in (("prob", ns["prob_map_mean"], prob), ("pred", ns["pred"], pred), ("out_feat", ns["out_feat"], of)):
        assert a.dtype == b.dtype and a.shape == b.shape and np.array_equal(a, b), name
    np.savez_compressed(f"{OUT}/tail.npz", input_sha=sha(logits, feat, inverse), prob_sha=sha(ns["prob_map_mean"]), prob_head=ns["prob_map_mean"][:64],
                        pred=ns["pred"].astype(np.int16), out_feat_head=ns["out_feat"][:64], out_feat_sha=sha(ns["out_feat"]))
    print(f"tail.npz: reference prob_inference.py:99-118 == oracle tta_tail / outfeat_mean on {reps} views x {len(pred)} points")


def gen_shard():
    """The reference's own score-mode data loader (dataset/sk_dataloader.py:185-198, unmodified) on a tree of empty .bin
    files: which frames rank r of G gets == lidal_b200.pipeline.frame_shard (the split every sharded driver uses)."""
    import torch.distributed as dist
    from dataset.sk_dataloader import SK_Dataloader
    from lidal_b200 import pipeline
    cases = [(1000, 1), (1000, 2), (1000, 4), (1000, 8), (1003, 8), (10, 4), (7, 8), (25, 3)]
    cwd = os.getcwd()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT="29653")
    dist.init_process_group("gloo", rank=0, world_size=1)           # SK_Dataloader.__init__ calls dist.barrier() when gpu_num > 1
    rows = []
    try:
        for n, world in cases:
            with tempfile.TemporaryDirectory() as d:
                os.makedirs(f"{d}/Processing_files")
                per = [n - n // 2, n // 2]                          # two of the train sequences hold the frames
                for seq_id, k in zip(("00", "01"), per):
                    os.makedirs(f"{d}/Semantic_kitti/dataset/sequences/{seq_id}/velodyne")
                    for i in range(k):
                        open(f"{d}/Semantic_kitti/dataset/sequences/{seq_id}/velodyne/{i:06d}.bin", "wb").close()
                os.chdir(d)
                try:
                    with redirect_stdout(io.StringIO()):
                        every = list(SK_Dataloader(gpu_num=1, gpu_rank=0, num_workers=0).score_data_loader(inf_reps=8).dataset.lidar_files[::8])
                        assert len(every) == n
                        for rank in range(world):
                            files = list(SK_Dataloader(gpu_num=world, gpu_rank=rank, num_workers=0).score_data_loader(inf_reps=8).dataset.lidar_files[::8])
                            got = [every.index(f) for f in files]
                            mine = pipeline.frame_shard(n, world, rank)
                            assert got == list(mine), (n, world, rank)
                            rows.append((n, world, rank, mine.start, mine.stop))
                finally:
                    os.chdir(cwd)
    finally:
        dist.destroy_process_group()
    np.savez_compressed(f"{OUT}/shard.npz", rows=np.array(rows, dtype=np.int64))
    print(f"shard.npz: reference score_data_loader split == pipeline.frame_shard on {len(cases)} (frames, ranks) cases, {len(rows)} shards")


def gen_cli():
    """The reference's own command line, `python -m score.sv_level.LiDAL --dataset_name SK --model_name SPVCNN --r_id 1`
    (README.md:115), run UNMODIFIED as a subprocess in a synthetic Processing_files/ tree (synth.write_scoring_tree; two of
    the ten SemanticKITTI train sequences populated, the others empty -- the reference handles that), against
    lidal_b200.score.cli.run on the pinned oracle: the sv_flag/*.npy files and the cached sv_pnums / sv_centers it writes
    must be byte-identical.  The concatenated outputs are committed as cli.npz."""
    import glob
    import subprocess
    from lidal_b200.score import cli
    params = dict(model_name="SPVCNN", lengths={"00": 26, "01": 25}, max_points=300, n_cls=19, seed=40)
    with tempfile.TemporaryDirectory() as d_ref, tempfile.TemporaryDirectory() as d_mine:
        n_regions = synth.write_scoring_tree(d_ref, **params)
        assert synth.write_scoring_tree(d_mine, **params) == n_regions
        env = dict(os.environ, PYTHONPATH=os.pathsep.join([REF, os.path.join(ROOT, "oracle", "_stubs")]))
        out = subprocess.run([sys.executable, "-m", "score.sv_level.LiDAL", "--dataset_name", "SK", "--model_name", "SPVCNN", "--r_id", "1"],
                             cwd=d_ref, env=env, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, timeout=1800)
        assert out.returncode == 0, out.stderr[-3000:]
        flags, paths = cli.run("SK", "SPVCNN", 1, root=d_mine, score_dataset=orc.score_dataset, select_regions=orc.select_regions,
                               verbose=False)
        rel = lambda root: sorted(os.path.relpath(p, root) for p in glob.glob(f"{root}/Processing_files/SK/sv_flag/KMeans/SPVCNN/LiDAL/1r/*/*.npy"))  # noqa: E731
        assert rel(d_ref) == rel(d_mine) and len(rel(d_ref)) == sum(params["lengths"].values())
        per_frame = []
        for r in rel(d_ref):
            a, b = np.load(os.path.join(d_ref, r)), np.load(os.path.join(d_mine, r))
            assert a.dtype == b.dtype and np.array_equal(a, b), r
            per_frame.append(a)
        stats = {}
        for name in ("sv_pnums", "sv_centers"):
            a = np.load(f"{d_ref}/Processing_files/SK/super_voxel/KMeans/{name}.npy")
            b = np.load(f"{d_mine}/Processing_files/SK/super_voxel/KMeans/{name}.npy")
            assert a.dtype == b.dtype and np.array_equal(a, b), name
            stats[name] = a
        for seq_id in cli.SK_TRAIN_SPLIT:                      # the reference creates every sequence's output folder (LiDAL.py:157-159)
            assert os.path.isdir(f"{d_ref}/Processing_files/SK/sv_flag/KMeans/SPVCNN/LiDAL/1r/{seq_id}")
            assert os.path.isdir(f"{d_mine}/Processing_files/SK/sv_flag/KMeans/SPVCNN/LiDAL/1r/{seq_id}")
    out_flags = np.concatenate(per_frame)
    assert np.array_equal(out_flags, flags)
    np.savez_compressed(f"{OUT}/cli.npz", params=np.array(list(params.items()), dtype=object), n_regions=n_regions,
                        flags=out_flags, frame_sizes=np.array([len(a) for a in per_frame]), **stats, allow_pickle=True)
    print(f"cli.npz: reference `python -m score.sv_level.LiDAL` == lidal_b200.score.cli.run on the oracle; {n_regions} regions, "
          f"labelled {int((out_flags == 1).sum())}, pseudo {int((out_flags == 2).sum())}")


if __name__ == "__main__":
    which = sys.argv[1:] or ["hash", "scoring", "selection", "nets", "voxelizer", "extra", "cli", "shard", "tail"]
    for w in which:
        globals()[f"gen_{w}"]()
