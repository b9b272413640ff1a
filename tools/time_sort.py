"""Times lb_sort_pairs and lb_kmap_sort_by_mask (A/B via LIDAL_SORT_3PHASE / LIDAL_MASK_KEY_BITS env)."""
import os, sys
sys.path[:0] = [os.getcwd()]
import torch
from lidal_b200 import _lib as L

def timed(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1000

for n, bits in ((10_000, 27), (131_000, 39), (900_000, 27), (900_000, 24), (4_000_000, 40)):
    keys = torch.randint(0, 2 ** bits, (n,), dtype=torch.int64).cuda()
    vals = torch.arange(n, dtype=torch.int32).cuda()
    nbytes = L.lib().lb_sort_pairs_ws_bytes(n)
    ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    k2, v2 = keys.clone(), vals.clone()
    us = timed(lambda: L.check(L.lib().lb_sort_pairs(L.ptr(k2), L.ptr(v2), n, bits, L.ptr(ws), nbytes, L.stream())))
    print(f"sort_pairs n={n} bits={bits}: {us:.1f} us")
