import sys, os
sys.path[:0] = [os.getcwd()]
import numpy as np, torch
import lidal_b200.compat as ts
from lidal_b200 import synth, engine
F = ts.nn.functional
c, f, _ = synth.scan_batch(seed=17, kind="SK", batch=8)
coords = torch.from_numpy(c).cuda()
m_s = engine.Maps(coords)
engine.SORT_MAPS = False
m_u = engine.Maps(coords)
def active(nbr, n):
    k = nbr.shape[0]
    nt = (n + 127) // 128
    v = torch.zeros((k, nt * 128), dtype=torch.bool, device="cuda"); v[:, :n] = nbr[:, :n] >= 0
    return v.view(k, nt, 128).any(2).float().mean().item()
for lvl in range(5):
    print("lvl", lvl, "n", m_s.n[lvl], "k3 active frac sorted", round(active(m_s.nbr3[lvl][0], m_s.n[lvl]), 3), "unsorted", round(active(m_u.nbr3[lvl], m_u.n[lvl]), 3))
def timeit(fn, reps=20):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
g = torch.Generator().manual_seed(0)
for lvl, cin, cout in ((0, 96, 96), (0, 32, 32), (1, 96, 96), (2, 128, 128), (3, 256, 256), (3, 384, 256), (4, 256, 256)):
    n = m_s.n[lvl]
    x = torch.randn(n, cin, generator=g).cuda().bfloat16()
    conv = engine._Conv((torch.randn(27, cin, cout, generator=g) * 0.05).cuda(), None, relu=True)
    out = torch.empty(n, cout, dtype=torch.bfloat16, device="cuda")
    ts_ = timeit(lambda: conv(x, m_s.nbr3[lvl], n, out=out))
    tu = timeit(lambda: conv(x, m_u.nbr3[lvl], n, out=out))
    pairs = int((m_u.nbr3[lvl] >= 0).sum())
    print(f"lvl{lvl} {cin}->{cout} n={n}: sorted {ts_:.3f} ms ({2*pairs*cin*cout/ts_/1e9:.0f} alg TF/s)  unsorted {tu:.3f} ms")
