"""One SPVCNN engine step on a batch-8 SK input (target for ncu -k filters; no timing here)."""
import sys, os
sys.path[:0] = [os.getcwd()]
import torch
import lidal_b200.compat as ts
from lidal_b200 import synth, engine
from lidal_b200.network import SPVCNN, seeded_state_dict
c, f, _ = synth.scan_batch(seed=17, kind="SK", batch=8)
coords, feats = torch.from_numpy(c).cuda(), torch.from_numpy(f).cuda()
model = SPVCNN(19, ts); model.load_state_dict(seeded_state_dict(model.state_dict())); model = model.cuda().eval()
eng = engine.InferenceEngine(model)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    eng(coords, feats)
torch.cuda.synchronize()
