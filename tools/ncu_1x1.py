"""A few 1x1 (identity-row) conv launches of the SPVCNN point branch for ncu captures: 32->256 with residual (relu first)."""
import sys, os
sys.path[:0] = [os.getcwd()]
import torch
from lidal_b200 import engine
n = 766073
g = torch.Generator().manual_seed(0)
for cin, cout in ((32, 256), (128, 96)):
    x = torch.randn(n, cin, generator=g).cuda().bfloat16()
    res = torch.randn(n, cout, generator=g).cuda().bfloat16()
    bn = torch.nn.BatchNorm1d(cout).cuda().eval()
    conv = engine._Conv((torch.randn(1, cin, cout, generator=g) * 0.05).cuda(), bn, relu=True)
    out = torch.empty(n, cout, dtype=torch.bfloat16, device="cuda")
    for _ in range(2):
        conv(x, None, n, out=out, residual=res, relu_first=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        conv(x, None, n, out=out, residual=res, relu_first=True)
    e1.record(); torch.cuda.synchronize()
    print(f"1x1 {cin}->{cout} n={n}: {e0.elapsed_time(e1)/10:.3f} ms", flush=True)
