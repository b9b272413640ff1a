import sys, os
sys.path[:0] = [os.getcwd(), os.path.join(os.getcwd(), "oracle")]
import numpy as np, torch
import torchsparse as oracle_ts
import lidal_b200.compat as ts
from lidal_b200 import synth
from lidal_b200.network import MinkUNet, seeded_state_dict
Fo = oracle_ts.nn.functional
raw = synth.raycast_scan(42, "NU"); rs = np.random.RandomState(9)
coords, feats, inv = synth.collate_views([synth.score_transform(raw[::4], rs), synth.score_transform(raw[1::4], rs)])
sel = np.arange(coords.shape[0]) % 3 == 0
coords, feats = np.ascontiguousarray(coords[sel]), np.ascontiguousarray(feats[sel])
labels = torch.from_numpy(np.random.default_rng(0).integers(0, 19, coords.shape[0])); labels[::11] = 255
acts = {}
def run(be, dev, tag):
    model = MinkUNet(19, be); model.load_state_dict(seeded_state_dict(model.state_dict()), strict=True); model = model.to(dev).train()
    hooks = []
    for name, mod in model.named_modules():
        if isinstance(mod, be.nn.Conv3d):
            def hook(m, i, o, name=name):
                o.F.retain_grad(); acts[(tag, name)] = o.F
            hooks.append(mod.register_forward_hook(hook))
    logits, _ = model(be.SparseTensor(torch.from_numpy(feats).to(dev), torch.from_numpy(coords).to(dev)))
    loss = torch.nn.functional.cross_entropy(logits, labels.to(dev), ignore_index=255); loss.backward()
    return float(loss), {k: p.grad.detach().cpu().double() for k, p in model.named_parameters() if k.endswith("kernel")}, model
lg, gg, mg = run(ts, "cuda", "g")
Fo.EMULATE_16BIT = torch.bfloat16
lo, go, mo = run(oracle_ts, "cpu", "o")
names = [n for (t, n) in acts if t == "g"]
print("layer | fwd act err | act-grad err | kernel-grad err")
for n in names:
    a, b = acts[("g", n)], acts[("o", n)]
    # coarse-level row order equals (both sorted by b,x,y,z in compat/oracle)
    fe = float((a.detach().cpu().double() - b.detach().double()).norm() / b.detach().double().norm())
    ge = float((a.grad.cpu().double() - b.grad.double()).norm() / b.grad.double().norm())
    ke = float((gg[n + ".kernel"] - go[n + ".kernel"]).norm() / go[n + ".kernel"].norm())
    print(f"{n:28s} {fe:.2e} {ge:.2e} {ke:.2e}")
