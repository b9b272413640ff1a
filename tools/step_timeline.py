"""Warm-cache section timeline of one engine step (CUDA events), for deciding what to optimise next."""
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
marks = []
def mark(name):
    e = torch.cuda.Event(enable_timing=True); e.record(); marks.append((name, e))
# wrap pieces
orig_iv, orig_maps = eng._initial_voxelize, engine.Maps
def iv(*a):
    mark("start"); r = orig_iv(*a); mark("initial_voxelize"); return r
eng._initial_voxelize = iv
class M2(orig_maps):
    def __init__(self, coords):
        super().__init__(coords); mark("maps")
engine.Maps = M2
for name in ("_corner_query", "_cell_query", "_devox", "_vox", "_cast", "_pad8"):
    fn = getattr(eng, name)
    def wrap(fn=fn, name=name):
        def w(*a, **k):
            mark("convs+other"); r = fn(*a, **k); mark(name); return r
        return w
    setattr(eng, name, wrap())
for it in range(3):
    marks.clear()
    out = eng(coords, feats); mark("convs+other")
    torch.cuda.synchronize()
agg = {}
for (n0, e0), (n1, e1) in zip(marks[:-1], marks[1:]):
    agg[n1] = agg.get(n1, 0.0) + e0.elapsed_time(e1)
tot = marks[0][1].elapsed_time(marks[-1][1])
for k, v in sorted(agg.items(), key=lambda kv: -kv[1]):
    print(f"{k:20s} {v:7.3f} ms  {100*v/tot:5.1f}%")
print("total", tot)
