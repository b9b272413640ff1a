"""Times single lb_conv_fwd launches (1x1 layers of the SPVCNN step) under the tile / epilogue variants."""
import os, sys
sys.path[:0] = [os.getcwd()]
import torch
import lidal_b200.compat as ts
from lidal_b200 import _lib as L
F = ts.nn.functional
n = 766_073
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def timed(fn, reps=10):
    for _ in range(2): fn()
    tot = 0.0
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        tot += e0.elapsed_time(e1)
    return tot / reps * 1000

g = torch.Generator().manual_seed(0)
SHAPES = ((32, 256, torch.bfloat16, False), (256, 128, torch.bfloat16, False), (128, 96, torch.bfloat16, False),
                            (128, 96, torch.bfloat16, True), (96, 32, torch.float32, False), (64, 64, torch.bfloat16, False))
ONE = len(sys.argv) > 2          # `time_layer.py CIN COUT`: one launch of that shape (for ncu)
if ONE:
    SHAPES = tuple(s for s in SHAPES if s[0] == int(sys.argv[1]) and s[1] == int(sys.argv[2]))[:1]
for cin, cout, odt, res in SHAPES:
    x = torch.randn(n, cin, generator=g).cuda().bfloat16()
    w = F.pack_weight((torch.randn(1, cin, cout, generator=g) * 0.1).cuda(), torch.bfloat16)
    sc, sh = torch.ones(cout, device="cuda"), torch.zeros(cout, device="cuda")
    r = torch.randn(n, cout, generator=g).cuda().bfloat16() if res else None
    out = torch.empty(n, cout, dtype=odt, device="cuda")
    byt = n * cin * 2 + n * cout * out.element_size() + (n * cout * 2 if res else 0)
    if ONE:
        F.conv_forward(x, w, None, n, scale=sc, shift=sh, residual=r, relu=True, out=out); torch.cuda.synchronize(); break
    line = f"1x1 {cin:3d}->{cout:3d} {str(odt)[6:]:8s} res={int(res)} ({byt/1e6:5.0f} MB): "
    for name, fl in (("default", 0), ("tile128", L.LB_CONV_TILE128), ("nostage", L.LB_CONV_NO_STAGED),
                     ("tile128+nostage", L.LB_CONV_TILE128 | L.LB_CONV_NO_STAGED)):
        us = timed(lambda: F.conv_forward(x, w, None, n, scale=sc, shift=sh, residual=r, relu=True, out=out, extra_flags=fl))
        line += f" {name} {us:6.1f} us ({byt/us/1e6:4.2f} TB/s)"
    print(line)
