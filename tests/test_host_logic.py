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
        d = np.sqrt(np.square(centers[i] - centers).sum(1))
        near = np.where((d < radius) & (np.arange(len(centers)) != i))[0]
        rows.append(near)
        ptr.append(ptr[-1] + len(near))
    return ptr, np.concatenate(rows)


def _walk_both_passes(flags, d, e, pn, c, tpn, argsort, index="grid"):
    cells = score.region_cells(c, 5.0)
    csr = _cpu_pairs(c, np.float32(5.0)) if index == "csr" else None
    new_index = (lambda: score._CsrIndex(*csr)) if csr else (lambda: score._GridIndex(c, cells, 5.0))
    limit = round(0.01 * int(tpn))
    ids = np.where(flags == 0)[0]
    order = argsort(d[ids])
    score._greedy_walk(order[::-1], ids, d, e, pn, new_index(), flags, 1, limit, True, False)
    ids = np.where(flags == 0)[0]
    order = argsort(d[ids])
    flags[flags == 2] = 0
    score._greedy_walk(order, ids, d, e, pn, new_index(), flags, 2, limit, False, True)
    return flags


def test_greedy_walk_replays_reference_selection(golden):
    """The host replay (set-iteration-order exact, added regions indexed by a 5 m grid) == reference golden flags."""
    g = golden["selection"]
    for d_key, tpn_key, out_key in (("sv_interds", "tight_tpn", "tight_out"), ("loose_interds", "loose_tpn", "loose_out")):
        flags = _walk_both_passes(g["sv_flags"].astype(int), g[d_key], g["sv_interes"], g["sv_pnums"],
                                  np.ascontiguousarray(g["sv_centers"], dtype=np.float32), g[tpn_key], np.argsort)
        assert np.array_equal(flags, g[out_key]), d_key


def test_greedy_walk_vs_oracle_on_a_sequence_shaped_region_set():
    """2,000 regions laid out like a driving sequence (20 sectors per frame, ego moving 1 m / frame): many candidates have
    several added regions within 5 m, so the set-order rule and the swap branch are exercised; == oracle selection."""
    import lidal_scoring as orc
    sv_flags, d, e, pn, c = synth.region_table(100, seed=3)
    tpn = int(pn.sum()) * 20            # budget = 20 % of the points: binds in both passes
    want = orc.select_regions(sv_flags.copy(), d, e, pn, c, tpn)
    got = _walk_both_passes(sv_flags.astype(int), d, e, pn, c, tpn, np.argsort)
    assert np.array_equal(got, want)
    got = _walk_both_passes(sv_flags.astype(int), d, e, pn, c, tpn, np.argsort, index="csr")
    assert np.array_equal(got, want)
    assert (got == 1).sum() > 20 and (got == 2).sum() > 20


def test_synthetic_shapes():
    raw = synth.raycast_scan(1, "SK")
    assert 110_000 < raw.shape[0] < 140_000 and raw.dtype == np.float32
    c, f, inv = synth.tta_batch(raw[::8], seed=3, inf_reps=8)
    assert c.dtype == np.int32 and c.shape[1] == 4 and f.shape == (c.shape[0], 4)
    assert inv.shape[0] == 8 * raw[::8].shape[0] and inv.max() == c.shape[0] - 1
    assert c[:, :3].min() >= 0 and c[:, :3].max() < 8192 and set(np.unique(c[:, 3])) == set(range(8))


def test_set_order_model_matches_the_interpreter():
    """_SetOrder replays CPython's set table for small ints: same iteration order as a real set after random add / remove
    sequences crossing several resizes (the check select_regions itself performs before trusting the model)."""
    assert score._set_model() is not None
    rng = np.random.RandomState(7)
    real, model, members = set(), score._SetOrder(score._set_model()), []
    for step in range(20000):
        if members and rng.rand() < 0.45:
            k = members.pop(rng.randint(len(members)))
            real.remove(k)
            model.remove(k)
        else:
            k = int(rng.randint(0, 200000))
            if k not in real:
                members.append(k)
            real.add(k)
            model.add(k)
    assert list(real) == list(model) and len(model) == len(real)
    slots = [model.slot[k] for k in real]
    assert slots == sorted(slots)


def _native_both_passes(flags, d, e, pn, c, tpn):
    ptr, nbr = _cpu_pairs(c, np.float32(5.0))
    limit = round(0.01 * int(tpn))
    ids = np.where(flags == 0)[0]
    score.native_walk(ids[np.argsort(d[ids])[::-1]], d, e, pn, ptr, nbr, flags, 1, limit, True, False, score._set_model())
    ids = np.where(flags == 0)[0]
    order = np.argsort(d[ids])
    flags[flags == 2] = 0
    score.native_walk(ids[order], d, e, pn, ptr, nbr, flags, 2, limit, False, True, score._set_model())
    return flags


def test_native_walk_equals_python_replay_and_golden(golden):
    """lb_select_walk (host C++, modelled set order) == the Python replay on a real set == the reference's golden flags."""
    g = golden["selection"]
    for d_key, tpn_key, out_key in (("sv_interds", "tight_tpn", "tight_out"), ("loose_interds", "loose_tpn", "loose_out")):
        flags = _native_both_passes(g["sv_flags"].astype(int), g[d_key], g["sv_interes"], g["sv_pnums"],
                                    np.ascontiguousarray(g["sv_centers"], dtype=np.float32), g[tpn_key])
        assert np.array_equal(flags, g[out_key]), d_key
    for seed, frames, frac in ((3, 100, 20), (5, 150, 8), (9, 60, 60)):
        sv_flags, d, e, pn, c = synth.region_table(frames, seed=seed)
        tpn = int(pn.sum()) * frac
        want = _walk_both_passes(sv_flags.astype(int), d, e, pn, c, tpn, np.argsort, index="csr")
        saved, score._SET_MODEL[:] = list(score._SET_MODEL), [None]          # force the real-set scan in the Python replay
        try:
            scan = _walk_both_passes(sv_flags.astype(int), d, e, pn, c, tpn, np.argsort, index="grid")
        finally:
            score._SET_MODEL[:] = saved
        got = _native_both_passes(sv_flags.astype(int), d, e, pn, c, tpn)
        assert np.array_equal(want, scan) and np.array_equal(got, want), seed
        assert (got == 1).sum() > 0 and (got == 2).sum() > 0


def test_accelerate_keeps_the_module_contract(oracle_ts, small_scan):
    """engine.accelerate(model) must leave the object a caller of the reference holds intact: same class, same state_dict
    keys, and the original (autograd-correct) forward whenever the module trains or autograd is on.  (The fused eval path
    itself needs the GPU: tests/test_gpu_engine.py.)"""
    import torch
    from lidal_b200.engine import accelerate
    from lidal_b200.network import MinkUNet, seeded_state_dict
    coords, feats, _ = small_scan
    model = MinkUNet(19, oracle_ts)
    model.load_state_dict(seeded_state_dict(model.state_dict()), strict=True)
    model.eval()
    keys = list(model.state_dict().keys())
    x = lambda: oracle_ts.SparseTensor(torch.from_numpy(feats[:400]), torch.from_numpy(coords[:400]))     # noqa: E731
    want = model(x())[0]                                               # autograd enabled: the module-by-module forward
    same = accelerate(model)
    assert same is model and type(model) is MinkUNet and accelerate(model) is model                      # in place, idempotent
    assert list(model.state_dict().keys()) == keys
    got = model(x())[0]
    assert got.requires_grad and torch.equal(got, want)                # grad mode: original forward, bit for bit
    model.train()
    assert model(x())[0].requires_grad
    model.eval()
    model.lidal_engine_invalidate()
    wrapped = torch.nn.Sequential()                                    # a DDP-like wrapper: `.module` is unwrapped
    wrapped.module = MinkUNet(19, oracle_ts).eval()
    assert accelerate(wrapped) is wrapped and wrapped.module._lidal_accelerated
