"""Run a few representative sparse-conv launches once each (for `ncu --metrics ...` memory-counter captures)."""
import sys, os
sys.path[:0] = [os.getcwd()]
import torch
import lidal_b200.compat as ts
from lidal_b200 import synth, engine
c, f, _ = synth.scan_batch(seed=17, kind="SK", batch=8)
coords = torch.from_numpy(c).cuda()
if "--lex" not in sys.argv:     # SPVCNN's internal voxel order: ascending hash (random in space)
    h = ts.nn.functional.sphash(coords)
    coords = coords[torch.argsort(h)].contiguous()
m = engine.Maps(coords)
g = torch.Generator().manual_seed(0)
for lvl, cin, cout in ((0, 96, 96), (0, 32, 32), (1, 96, 96), (2, 128, 128), (3, 256, 256), (4, 256, 256)):
    n = m.n[lvl]
    x = engine._alloc(n, cin, torch.bfloat16, "cuda")
    x.copy_(torch.randn(n, cin, generator=g).cuda())
    conv = engine._Conv((torch.randn(27, cin, cout, generator=g) * 0.05).cuda(), None, relu=True)
    out = torch.empty(n, cout, dtype=torch.bfloat16, device="cuda")
    for _ in range(2):
        conv(x, m.nbr3[lvl], n, out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        conv(x, m.nbr3[lvl], n, out=out)
    e1.record(); torch.cuda.synchronize()
    print(f"lvl{lvl} {cin}->{cout} n={n}: {e0.elapsed_time(e1)/10:.3f} ms", flush=True)
