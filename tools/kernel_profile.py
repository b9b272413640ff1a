"""Warm-cache per-kernel device time of one engine step (torch.profiler / CUPTI)."""
import sys, os
sys.path[:0] = [os.getcwd()]
import torch
from torch.profiler import profile, ProfilerActivity
import lidal_b200.compat as ts
from lidal_b200 import synth, engine
from lidal_b200.network import SPVCNN, seeded_state_dict
c, f, _ = synth.scan_batch(seed=17, kind="SK", batch=8)
coords, feats = torch.from_numpy(c).cuda(), torch.from_numpy(f).cuda()
model = SPVCNN(19, ts); model.load_state_dict(seeded_state_dict(model.state_dict())); model = model.cuda().eval()
eng = engine.InferenceEngine(model)
for _ in range(3): eng(coords, feats)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3): eng(coords, feats)
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total / 3.0, e.count // 3) for e in prof.key_averages() if e.device_time_total > 0]
tot = sum(r[1] for r in rows)
print(f"device time per step: {tot/1e3:.3f} ms")
for k, t, n in sorted(rows, key=lambda r: -r[1])[:28]:
    print(f"{t:9.1f} us {100*t/tot:5.1f}%  x{n:3d}  {k[:90]}")
