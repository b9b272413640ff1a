"""Config 4 timing: MinkUNet training fwd + bwd (+ Adam step), synthetic SK-shaped, batch 2 per GPU, through lidal_b200.compat."""
import sys, os
sys.path[:0] = [os.getcwd()]
import numpy as np, torch
import lidal_b200.compat as ts
from lidal_b200 import synth
from lidal_b200.network import MinkUNet, seeded_state_dict
c, f, _ = synth.scan_batch(seed=5, kind="SK", batch=2)
coords, feats = torch.from_numpy(c).cuda(), torch.from_numpy(f).cuda()
labels = torch.randint(0, 19, (coords.shape[0],), device="cuda")
model = MinkUNet(19, ts); model.load_state_dict(seeded_state_dict(model.state_dict())); model = model.cuda().train()
opt = torch.optim.Adam(model.parameters())
def step():
    opt.zero_grad()
    logits, _ = model(ts.SparseTensor(feats, coords))
    loss = torch.nn.functional.cross_entropy(logits, labels, ignore_index=255)
    loss.backward(); opt.step()
    return loss
for _ in range(3): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): loss = step()
e1.record(); torch.cuda.synchronize()
print(f"MinkUNet train step (fwd+bwd+Adam), batch 2 SK scans ({coords.shape[0]} voxels): {e0.elapsed_time(e1)/10:.1f} ms/step, loss {float(loss):.4f}")
