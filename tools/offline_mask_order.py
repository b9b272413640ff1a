"""CPU-only study of row orderings for tile-level offset skipping: how many (tile, offset) blocks stay active.
Usage: python tools/offline_mask_order.py [level]"""
import os, sys
sys.path[:0] = [os.getcwd()]
import numpy as np
from lidal_b200 import synth

def masks_of(coords, stride):
    """bit k of mask[i] = neighbour at offset k present (27 offsets, any fixed order)."""
    c = coords.astype(np.int64)
    R = 1 << 14
    def key(x, y, z, b): return ((b * R + (x + 64)) * R + (y + 64)) * R + (z + 64)
    keys = key(c[:, 0], c[:, 1], c[:, 2], c[:, 3])
    sk = np.sort(keys)
    m = np.zeros(len(c), dtype=np.uint32)
    k = 0
    for dx in (-1, 0, 1):
        for dy in (-1, 0, 1):
            for dz in (-1, 0, 1):
                q = key(c[:, 0] + dx * stride, c[:, 1] + dy * stride, c[:, 2] + dz * stride, c[:, 3])
                pos = np.searchsorted(sk, q)
                pos[pos >= len(sk)] = len(sk) - 1
                m |= (sk[pos] == q).astype(np.uint32) << np.uint32(k)
                k += 1
    return m

def active_blocks(m_sorted, tile):
    n = len(m_sorted)
    nt = (n + tile - 1) // tile
    pad = np.zeros(nt * tile, dtype=np.uint32); pad[:n] = m_sorted
    u = np.bitwise_or.reduce(pad.reshape(nt, tile), axis=1)
    pc = np.array([bin(int(x)).count("1") for x in u])
    return pc.sum(), pc.mean(), nt

def freq_key(m):
    cnt = np.array([int(((m >> np.uint32(k)) & 1).sum()) for k in range(27)])
    order = np.argsort(cnt, kind="stable")           # ascending frequency; rank 0 = rarest -> MSB
    key = np.zeros(len(m), dtype=np.uint64)
    for rank, k in enumerate(order):
        key |= (((m >> np.uint32(k)) & 1).astype(np.uint64)) << np.uint64(26 - rank)
    return key, cnt

def downsample(coords, s):
    c = coords.copy(); c[:, :3] = c[:, :3] // (2 * s) * (2 * s)
    return np.unique(c, axis=0)

if __name__ == "__main__":
    lvl = int(sys.argv[1]) if len(sys.argv) > 1 else 0
    c, f, _ = synth.scan_batch(seed=17, kind="SK", batch=8)
    coords = np.unique(c, axis=0)
    for l in range(lvl): coords = downsample(coords, 2 ** l)
    s = 2 ** lvl
    m = masks_of(coords, s)
    pc_row = np.array([bin(int(x)).count("1") for x in m[:200000]]).mean()
    print(f"level {lvl}: n={len(m)} mean offsets/row={pc_row:.2f}")
    key, cnt = freq_key(m)
    orders = {
        "unsorted": np.arange(len(m)),
        "freq-key >>3 (current)": np.argsort(key >> np.uint64(3), kind="stable"),
        "freq-key full": np.argsort(key, kind="stable"),
        "plain mask": np.argsort(m, kind="stable"),
    }
    pcs = np.array([bin(int(x)).count("1") for x in m]) if len(m) < 2_000_000 else None
    if pcs is not None:
        orders["popcount, then freq-key"] = np.lexsort((key, pcs))
    rows = np.arange(len(m), dtype=np.uint64)
    for shift, drop in ((17, 3), (17, 6), (16, 6), (16, 7), (15, 7)):
        ck = ((rows >> np.uint64(shift)) << np.uint64(27)) | ((key >> np.uint64(drop)) << np.uint64(drop))
        orders[f"chunk 2^{shift} rows, key>>{drop}"] = np.argsort(ck, kind="stable")
    for name, o in orders.items():
        for tile in (128, 256):
            tot, mean, nt = active_blocks(m[o], tile)
            print(f"  {name:28s} tile {tile}: {mean:5.2f} active offsets/tile, rows*offsets executed = {tot * tile / len(m):6.2f} per row")
