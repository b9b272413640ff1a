"""Per-kernel device time of the LiDAL frame chain (torch.profiler / CUPTI, warm): voxelizer -> engine -> tail -> grid -> scoring."""
import sys, os
sys.path[:0] = [os.getcwd()]
import torch
from torch.profiler import profile, ProfilerActivity
import lidal_b200.compat as ts
from lidal_b200 import synth, engine, pipeline
from lidal_b200.network import SPVCNN, seeded_state_dict
n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
model = SPVCNN(19, ts); model.load_state_dict(seeded_state_dict(model.state_dict())); model = model.cuda().eval()
eng = engine.InferenceEngine(model)
seq = synth.GpuSequence(n, "SK", seed=3, device="cuda")
frames = {f: seq.frame(f) for f in range(n)}
src = lambda f: frames[f]
pipeline.run_sequence_sharded(eng, src, n, 19, n * 20)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    out = pipeline.run_sequence_sharded(eng, src, n, 19, n * 20)
    torch.cuda.synchronize()
tm = out[-1]
print({k: (round(v, 2) if isinstance(v, float) else v) for k, v in tm.items()})
rows = [(e.key, e.device_time_total / n, e.count / n) for e in prof.key_averages() if e.device_time_total > 0]
tot = sum(r[1] for r in rows)
print(f"device time per frame: {tot/1e3:.3f} ms")
for k, t, c in sorted(rows, key=lambda r: -r[1])[:45]:
    print(f"{t:9.1f} us {100*t/tot:5.1f}%  x{c:6.1f}  {k[:100]}")
