"""Host part of the global selection at BASELINE size, measured on the CPU alone (it IS host code): the native greedy walk
(`lb_select_walk`, csrc/select.cu) over CSR neighbour lists, for config 3 (20,000 regions) and config 5 (850 sequences x
40 frames x 20 regions = 680,000 regions), checked against the Python replay (`score._greedy_walk`) of LiDAL.py:242-325.

The 5 m pair lists come from scipy's cKDTree here (on the GPU they come from `lb_region_pairs`); candidates are filtered with
the reference's float32 expression `np.sqrt(np.square(a - b).sum()) < 5` (LiDAL.py:252-254).

    python tools/select_walk_scale.py [--sequences 850] [--frames 40] [--check]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def synth_regions(n_seq, n_frames, regions=20, seed=0):
    """Region centres the way a driving sequence produces them: per frame, `regions` azimuth sectors around an ego pose that
    advances 1 m per frame; sequence idx is offset by idx * 1000.0 (LiDAL.py:218).  float32, like the reference's arrays."""
    rng = np.random.default_rng(seed)
    ang = (np.arange(regions) + 0.5) * (2 * np.pi / regions)
    ring = np.stack([np.cos(ang), np.sin(ang), np.zeros(regions)], 1) * rng.uniform(6.0, 14.0, (regions, 1))
    c = np.zeros((n_seq, n_frames, regions, 3), np.float32)
    for s in range(n_seq):
        ego = np.stack([np.arange(n_frames, dtype=np.float64), 0.3 * np.sin(np.arange(n_frames) / 7.0 + s), np.zeros(n_frames)], 1)
        local = (ego[:, None, :] + ring[None] + rng.normal(0, 0.4, (n_frames, regions, 3))).astype(np.float32)
        c[s] = local + np.float32(0) + s * 1000.0
    n = n_seq * n_frames * regions
    d = rng.gamma(2.0, 1e-3, n).astype(np.float32)
    d[rng.random(n) < 0.02] = 0.0                                # regions without a single matched point
    e = rng.uniform(0.2, 2.5, n).astype(np.float32)
    pn = rng.integers(1500, 1800, n)
    return c.reshape(-1, 3), d, e, pn


def pairs_csr(centers, radius=5.0):
    from scipy.spatial import cKDTree
    tree = cKDTree(centers.astype(np.float64))
    pr = tree.query_pairs(radius * 1.001, output_type="ndarray")
    a, b = centers[pr[:, 0]], centers[pr[:, 1]]
    keep = np.sqrt(np.square(a - b).sum(1)) < np.float32(radius)               # float32, the reference's operation order
    pr = pr[keep]
    both = np.concatenate([pr, pr[:, ::-1]])
    both = both[np.lexsort((both[:, 1], both[:, 0]))]
    row_ptr = np.zeros(len(centers) + 1, np.int64)
    np.add.at(row_ptr, both[:, 0] + 1, 1)
    return np.cumsum(row_ptr), both[:, 1].astype(np.int32)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sequences", type=int, default=850)
    ap.add_argument("--frames", type=int, default=40)
    ap.add_argument("--train-point-num", type=int, default=976677792)            # LiDAL.py:135 (nuScenes)
    ap.add_argument("--check", action="store_true", help="also run the Python replay and compare flags")
    args = ap.parse_args()
    from lidal_b200 import score
    centers, d, e, pn = synth_regions(args.sequences, args.frames)
    n = len(d)
    n_frames = args.sequences * args.frames
    flags0 = np.zeros(n, int)
    lab = np.random.default_rng(5).choice(n_frames, max(1, n_frames // 100), replace=False)
    flags0.reshape(n_frames, -1)[lab] = 1
    t0 = time.perf_counter()
    row_ptr, nbr = pairs_csr(centers)
    t_pairs = time.perf_counter() - t0
    model = score._set_model()
    assert model is not None, "set-order model does not match this interpreter"
    d64, e64, pn64 = d.astype(np.float64), e.astype(np.float64), pn.astype(np.int64)
    limit = round(0.01 * args.train_point_num)

    def run(walk):
        flags = flags0.copy()
        ids = np.where(flags == 0)[0]
        order = np.argsort(d[ids])
        t = time.perf_counter()
        walk(flags, ids, order[::-1], 1, True, False)
        t1 = time.perf_counter() - t
        ids = np.where(flags == 0)[0]
        order = np.argsort(d[ids])
        flags[flags == 2] = 0
        t = time.perf_counter()
        walk(flags, ids, order, 2, False, True)
        return flags, t1, time.perf_counter() - t

    def native(flags, ids, order, value, higher, skip):
        score.native_walk(ids[order], d64, e64, pn64, row_ptr, nbr, flags, value, limit, higher, skip, model)

    def replay(flags, ids, order, value, higher, skip):
        score._greedy_walk(order, ids, d, e, pn, score._CsrIndex(row_ptr.tolist(), nbr), flags, value, limit, higher, skip)

    flags, t_al, t_sl = run(native)
    print(f"regions {n}  pairs {len(nbr)} ({len(nbr) / n:.1f} per region, cKDTree {t_pairs * 1e3:.0f} ms)  budget {limit} points")
    print(f"native walk: AL pass {t_al * 1e3:.1f} ms, pseudo-label pass {t_sl * 1e3:.1f} ms  -> labelled {(flags == 1).sum() - (flags0 == 1).sum()} "
          f"new regions, pseudo {(flags == 2).sum()}")
    if args.check:
        ref, r_al, r_sl = run(replay)
        print(f"python replay: {r_al * 1e3:.0f} + {r_sl * 1e3:.0f} ms  flags equal: {np.array_equal(ref, flags)}")
        assert np.array_equal(ref, flags)


if __name__ == "__main__":
    main()
