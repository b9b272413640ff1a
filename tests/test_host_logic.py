"""CPU: host-side logic of the scoring path (no device calls)."""
import numpy as np

from lidal_b200 import score, synth


def test_neighbour_ids_match_oracle():
    import lidal_scoring as orc
    for n in (25, 26, 40, 1000):
        for fid in (0, 1, 11, 12, 13, n // 2, n - 13, n - 12, n - 2, n - 1):
            assert score.neighbour_ids(fid, n) == orc.neighbour_ids(fid, n)


def _cpu_pairs(centers, radius):
    rows, ptr = [], [0]
    for i in range(len(centers)):
        d = np.sqrt(np.square(centers[i] - centers).sum(1, dtype=np.float32))
        near = np.where((d < radius) & (np.arange(len(centers)) != i))[0]
        rows.append(near)
        ptr.append(ptr[-1] + len(near))
    return np.array(ptr), np.concatenate(rows).astype(np.int64)


def test_greedy_walk_replays_reference_selection(golden):
    """The host replay (set-iteration-order exact) fed with CPU-built neighbour lists == reference golden flags."""
    g = golden["selection"]
    for d_key, tpn_key, out_key in (("sv_interds", "tight_tpn", "tight_out"), ("loose_interds", "loose_tpn", "loose_out")):
        flags = g["sv_flags"].astype(int)
        d, e, pn, c = g[d_key], g["sv_interes"], g["sv_pnums"], g["sv_centers"]
        ptr, idx = _cpu_pairs(c, np.float32(5.0))
        limit = round(0.01 * int(g[tpn_key]))
        ids = np.where(flags == 0)[0]
        order = np.argsort(d[ids], kind="stable")
        score._greedy_walk(order[::-1], ids, d, e, pn, ptr, idx, flags, 1, limit, True, False)
        ids = np.where(flags == 0)[0]
        order = np.argsort(d[ids], kind="stable")
        flags[flags == 2] = 0
        score._greedy_walk(order, ids, d, e, pn, ptr, idx, flags, 2, limit, False, True)
        assert np.array_equal(flags, g[out_key]), d_key


def test_synthetic_shapes():
    raw = synth.raycast_scan(1, "SK")
    assert 110_000 < raw.shape[0] < 140_000 and raw.dtype == np.float32
    c, f, inv = synth.tta_batch(raw[::8], seed=3, inf_reps=8)
    assert c.dtype == np.int32 and c.shape[1] == 4 and f.shape == (c.shape[0], 4)
    assert inv.shape[0] == 8 * raw[::8].shape[0] and inv.max() == c.shape[0] - 1
    assert c[:, :3].min() >= 0 and c[:, :3].max() < 8192 and set(np.unique(c[:, 3])) == set(range(8))
