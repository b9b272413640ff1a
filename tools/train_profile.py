"""Where does the host time of one MinkUNet training step go (config 4 shape)?  torch.profiler CPU + CUDA tables."""
import sys, os
sys.path[:0] = [os.getcwd()]
import numpy as np, torch
from torch.profiler import profile, ProfilerActivity
import lidal_b200.compat as ts
from lidal_b200 import synth
from lidal_b200.network import MinkUNet, seeded_state_dict
sets = [synth.scan_batch(seed=50 + s, kind="SK", batch=2) for s in range(2)]
data = [(torch.from_numpy(c).cuda(), torch.from_numpy(f).cuda(), torch.randint(0, 19, (c.shape[0],), device="cuda")) for c, f, _ in sets]
model = MinkUNet(19, ts); model.load_state_dict(seeded_state_dict(model.state_dict())); model = model.cuda().train()
opt = torch.optim.Adam(model.parameters())
def step(i):
    c, f, y = data[i % 2]
    opt.zero_grad()
    logits, _ = model(ts.SparseTensor(f, c))
    loss = torch.nn.functional.cross_entropy(logits, y, ignore_index=255)
    loss.backward(); opt.step()
    return loss
for i in range(4): step(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(10): step(i)
e1.record(); torch.cuda.synchronize()
print(f"train step: {e0.elapsed_time(e1)/10:.1f} ms")
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for i in range(4): step(i)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="self_cpu_time_total", row_limit=28, max_name_column_width=60))
print(prof.key_averages().table(sort_by="self_cuda_time_total", row_limit=22, max_name_column_width=60))
