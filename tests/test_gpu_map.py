"""GPU: hashing, hash query, downsample and kernel maps through the C ABI -- bit-exact against the oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ts():
    import lidal_b200.compat as ts
    return ts


def rand_coords(n, seed, batch=3, span=200):
    rng = np.random.default_rng(seed)
    c = np.concatenate([rng.integers(0, span, (n, 3)), rng.integers(0, batch, (n, 1))], 1).astype(np.int32)
    return np.unique(c, axis=0)


def test_sphash_golden(ts, golden):
    g = golden["hash"]
    got = ts.nn.functional.sphash(torch.from_numpy(g["coords"]).cuda())
    assert got.dtype == torch.int64 and np.array_equal(got.cpu().numpy(), g["hashes"])


@pytest.mark.parametrize("n", [0, 1, 31, 5000])
def test_sphash_and_kernel_hash_vs_oracle(ts, oracle_ts, n):
    c = torch.from_numpy(rand_coords(n, n)) if n else torch.zeros((0, 4), dtype=torch.int)
    F, Fo = ts.nn.functional, oracle_ts.nn.functional
    assert torch.equal(F.sphash(c.cuda()).cpu(), Fo.sphash(c))
    for ks, st in ((3, 1), (2, 2), (3, 4)):
        off = oracle_ts.nn.utils.get_kernel_offsets(ks, st)
        assert torch.equal(ts.nn.utils.get_kernel_offsets(ks, st, device="cuda").cpu(), off)
        assert torch.equal(F.sphash(c.cuda(), off.cuda()).cpu(), Fo.sphash(c, off))


def test_sphashquery_vs_oracle(ts, oracle_ts):
    F, Fo = ts.nn.functional, oracle_ts.nn.functional
    c = torch.from_numpy(rand_coords(20000, 1))
    ref = Fo.sphash(c)
    off = oracle_ts.nn.utils.get_kernel_offsets(3, 1)
    q = Fo.sphash(c[:7000], off)                                       # [27, 7000]: hits and misses
    want = Fo.sphashquery(q, ref)
    got = F.sphashquery(q.cuda(), ref.cuda())
    assert got.shape == q.shape and got.dtype == torch.int64
    assert torch.equal(got.cpu(), want)
    assert (want == -1).any() and (want >= 0).any()
    assert F.sphashquery(q[:0].cuda(), ref.cuda()).numel() == 0


def test_spcount(ts, oracle_ts):
    idx = torch.randint(-1, 50, (10000,), dtype=torch.int)
    assert torch.equal(ts.nn.functional.spcount(idx.cuda(), 50).cpu(), oracle_ts.nn.functional.spcount(idx, 50))


@pytest.mark.parametrize("ts_stride", [1, 2, 8])
def test_spdownsample_sorted_unique(ts, oracle_ts, ts_stride):
    c = rand_coords(30000, 7, batch=4, span=300)
    c[:, :3] = c[:, :3] // ts_stride * ts_stride
    c = torch.from_numpy(np.unique(c, axis=0))
    want = oracle_ts.nn.functional.spdownsample(c, 2, 2, ts_stride)
    got = ts.nn.functional.spdownsample(c.cuda(), 2, 2, ts_stride)
    assert torch.equal(got.cpu(), want)


def test_spdownsample_rejects_out_of_range(ts):
    from lidal_b200._lib import LidalError
    c = torch.tensor([[1, 2, 3, 0], [70000, 1, 1, 0]], dtype=torch.int).cuda()
    with pytest.raises(LidalError):
        ts.nn.functional.spdownsample(c, 2, 2, 1)


def test_kernel_maps_bit_exact(ts, oracle_ts, small_scan, golden):
    """All 9 maps of a MinkUNet forward: nbmaps / nbsizes / out coords identical to the oracle and the golden checksums."""
    import hashlib
    F, Fo = ts.nn.functional, oracle_ts.nn.functional
    coords = torch.from_numpy(small_scan[0])
    cur_o, cur_g = coords, coords.cuda()
    g = golden["nets"]
    for level in range(5):
        s = 2 ** level
        st = (s, s, s)
        nb_o, ns_o, sz_o, _, res_o = Fo.build_kernel_map(cur_o, st, (3, 3, 3), (1, 1, 1), (1, 1, 1))
        km = F.build_kernel_map(cur_g, st, (3, 3, 3), (1, 1, 1), (1, 1, 1))
        assert torch.equal(km.nbr.cpu().long(), res_o)
        assert torch.equal(km.nbmaps.cpu().long(), nb_o) and torch.equal(km.nbsizes.cpu().long(), ns_o)
        assert km[2] == sz_o
        tag = f"kmap_s{s}_k3_st1"
        assert hashlib.sha1(km.nbmaps.cpu().numpy().astype(np.int32).tobytes()).hexdigest() == str(g[tag + "_sha"])
        if level == 4:
            break
        nb_o, ns_o, sz_o, oc_o, res_o = Fo.build_kernel_map(cur_o, st, (2, 2, 2), (2, 2, 2), (1, 1, 1))
        km = F.build_kernel_map(cur_g, st, (2, 2, 2), (2, 2, 2), (1, 1, 1))
        assert torch.equal(km.out_coords.cpu(), oc_o)
        assert torch.equal(km.nbr.cpu().long(), res_o)
        assert torch.equal(km.nbmaps.cpu().long(), nb_o) and torch.equal(km.nbsizes.cpu().long(), ns_o)
        # transposed table is the per-offset inverse
        nt = km.nbr_t.cpu()
        k_idx, o_idx = torch.nonzero(res_o >= 0, as_tuple=True)
        assert torch.equal(nt[k_idx, res_o[k_idx, o_idx]].long(), o_idx)
        assert int((nt >= 0).sum()) == len(o_idx)
        cur_o, cur_g = oc_o, km.out_coords


def test_sort_pairs_stable_and_sorted():
    import ctypes as C
    from lidal_b200 import _lib as L
    n = 300_000
    keys = torch.randint(0, 2 ** 20, (n,), dtype=torch.int64).cuda() * 4096      # many duplicates, 32 bits used
    vals = torch.arange(n, dtype=torch.int32).cuda()
    k2, v2 = keys.clone(), vals.clone()
    nbytes = L.lib().lb_sort_pairs_ws_bytes(n)
    ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    L.check(L.lib().lb_sort_pairs(L.ptr(k2), L.ptr(v2), n, 32, L.ptr(ws), nbytes, L.stream()))
    want_k, want_i = torch.sort(keys.cpu(), stable=True)
    assert torch.equal(k2.cpu(), want_k) and torch.equal(v2.cpu().long(), want_i)


@pytest.mark.parametrize("n,bits,dup", [(1, 8, 1), (4097, 16, 1), (1_000_003, 27, 1), (6_000_000, 64, 1), (2_000_000, 40, 1 << 39)])
def test_sort_pairs_sizes(n, bits, dup):
    """Single-sweep radix sort (decoupled look-back): tile counts below and above what is resident at once, 1..8 passes,
    all-equal digits (dup = one huge multiplier leaves a single live bit)."""
    from lidal_b200 import _lib as L
    g = torch.Generator().manual_seed(n % 1000)
    if bits == 64:
        keys = torch.randint(-2 ** 63, 2 ** 63 - 1, (n,), dtype=torch.int64, generator=g)
    elif dup > 1:
        keys = torch.randint(0, 2, (n,), dtype=torch.int64, generator=g) * dup
    else:
        keys = torch.randint(0, 2 ** bits, (n,), dtype=torch.int64, generator=g)
    keys = keys.cuda()
    vals = torch.arange(n, dtype=torch.int32).cuda()
    k2, v2 = keys.clone(), vals.clone()
    nbytes = L.lib().lb_sort_pairs_ws_bytes(n)
    ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    for _ in range(2):          # second call reuses the dirty workspace
        k2.copy_(keys); v2.copy_(vals)
        L.check(L.lib().lb_sort_pairs(L.ptr(k2), L.ptr(v2), n, bits, L.ptr(ws), nbytes, L.stream()))
    # unsigned order: compare through a bias so torch's signed stable sort gives the same permutation
    biased = keys ^ (-2 ** 63) if bits == 64 else keys
    want_k, want_i = torch.sort(biased, stable=True)
    assert torch.equal(v2.long(), want_i)
    assert torch.equal(k2, keys[want_i])


def test_symmetric_kernel_map_query_matches_plain(ts, small_scan):
    """lb_kmap_query_sym (probe half the offsets, mirror the rest) == lb_kmap_query, bit for bit, at strides 1 and 2."""
    from lidal_b200 import _lib as L
    from lidal_b200 import engine
    F = ts.nn.functional
    coords = torch.from_numpy(small_scan[0]).cuda()
    for stride in (1, 2):
        c = coords if stride == 1 else F.spdownsample(coords, 2, 2, 1)
        n = c.shape[0]
        table = F._build_table(F.sphash(c))
        off = engine._offsets(3, stride, c.device)
        a = torch.empty((27, n), dtype=torch.int, device="cuda")
        b = torch.full((27, n + 5), 123, dtype=torch.int, device="cuda")
        L.check(L.lib().lb_kmap_query(L.ptr(table[0]), table[1], L.ptr(c), n, None, L.ptr(off), 27, L.ptr(a), L.stream()))
        L.check(L.lib().lb_kmap_query_sym(L.ptr(table[0]), table[1], L.ptr(c), n, L.ptr(off), 27, L.ptr(b), b.stride(0), L.stream()))
        assert torch.equal(a, b[:, :n])
        assert int((a >= 0).sum()) > n          # non-trivial map
        assert torch.equal(a[13].long().cpu(), torch.arange(n))


def test_level_counts_vs_numpy():
    """lb_level_counts: distinct parents at every coarser level in one pass (points or voxels, duplicates allowed) == numpy."""
    import ctypes as C
    from lidal_b200 import _lib as L
    rng = np.random.default_rng(5)
    for n in (0, 1, 777, 200000):
        c = np.concatenate([rng.integers(0, 300, (n, 3)), rng.integers(0, 4, (n, 1))], 1).astype(np.int32)
        if n > 10:
            c[n // 2:] = c[: n - n // 2]                      # duplicates
        ct = torch.from_numpy(c).cuda()
        counts = torch.full((4,), -7, dtype=torch.int32, device="cuda")
        nbytes = L.lib().lb_level_counts_ws_bytes(n, 4)
        ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
        L.check(L.lib().lb_level_counts(L.ptr(ct) if n else None, n, 4, L.ptr(counts), L.ptr(ws), nbytes, L.stream()))
        want = [len(np.unique(np.concatenate([c[:, :3] // (1 << l), c[:, 3:]], 1), axis=0)) if n else 0 for l in range(1, 5)]
        assert counts.cpu().tolist() == want
