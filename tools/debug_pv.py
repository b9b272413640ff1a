import sys, os
sys.path[:0] = [os.getcwd(), os.path.join(os.getcwd(), "oracle")]
import numpy as np, torch
import torchsparse as oracle_ts
import lidal_b200.compat as ts
from lidal_b200 import synth
raw = synth.raycast_scan(42, "NU"); rs = np.random.RandomState(9)
coords, feats, inv = synth.collate_views([synth.score_transform(raw[::4], rs), synth.score_transform(raw[1::4], rs)])
g = torch.Generator().manual_seed(0)
n = coords.shape[0]
f32 = torch.randn(n, 32, generator=g)
jitter = torch.rand(n, 3, generator=g) * 0.98
pc = torch.cat([torch.from_numpy(coords[:, :3]).float() + jitter, torch.from_numpy(coords[:, 3:]).float()], 1)
m = 5000
idx = torch.randint(-1, m, (n, 8), generator=g)
w = torch.rand(n, 8, generator=g)
vf = torch.randn(m, 32, generator=g)
a = oracle_ts.nn.functional.spdevoxelize(vf, idx, w)
b = ts.nn.functional.spdevoxelize(vf.cuda(), idx.cuda(), w.cuda()).cpu()
print("devox random:", (a - b).abs().max().item())
iq = torch.randint(-1, m, (8, n), generator=g)
for scale in (1, 4):
    wa = oracle_ts.nn.functional.calc_ti_weights(pc, iq, scale)
    wb = ts.nn.functional.calc_ti_weights(pc.cuda(), iq.cuda(), scale).cpu()
    d = (wa - wb).abs()
    print("ti_weights scale", scale, d.max().item(), "at", np.unravel_index(int(d.argmax()), d.shape), wa.flatten()[d.argmax()].item(), wb.flatten()[d.argmax()].item())
