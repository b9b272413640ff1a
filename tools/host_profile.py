"""Host-side cost of one engine step: wall time of the submit loop vs device time, and a cProfile of 10 steps."""
import os, sys, time, cProfile, pstats, io
sys.path[:0] = [os.getcwd()]
import torch
import bench
from lidal_b200.engine import StreamPipeline
dev = torch.device("cuda:0")
run, eng = bench.build_runner("spvcnn", "engine", dev)
batches = bench.make_batches(0)
resident = [(torch.from_numpy(c).to(dev), torch.from_numpy(f).to(dev)) for c, f, _ in batches]
sp = StreamPipeline(eng)
for i in range(9):
    sp.submit(*resident[i % 3], wait_main=False)
torch.cuda.synchronize()
# (1) host time of prepare / forward separately, GPU idle in between (pure host cost incl. the prepare round trips)
tp = tf = 0.0
for i in range(9):
    t0 = time.perf_counter(); pr = sp.prepare(*resident[i % 3], wait_main=False); t1 = time.perf_counter()
    torch.cuda.current_stream().wait_event(pr.ready); out = eng.forward(pr); t2 = time.perf_counter()
    torch.cuda.synchronize()
    tp += t1 - t0; tf += t2 - t1
print(f"host ms per step, GPU drained between steps: prepare {tp / 9 * 1e3:.2f} (includes its device round trips), forward launch loop {tf / 9 * 1e3:.2f}")
# (2) steady state
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); t0 = time.perf_counter()
for i in range(20):
    sp.submit(*resident[i % 3], wait_main=False)
t1 = time.perf_counter(); e1.record(); torch.cuda.synchronize()
print(f"steady state: host loop {(t1 - t0) / 20 * 1e3:.2f} ms/step, device {e0.elapsed_time(e1) / 20:.2f} ms/step")
# (3) device time of the two phases on their own
prs = [sp.prepare(*resident[i], wait_main=False) for i in range(3)]
torch.cuda.synchronize()
e0.record()
for i in range(20):
    eng.forward(prs[i % 3])
e1.record(); torch.cuda.synchronize()
print(f"forward only: device {e0.elapsed_time(e1) / 20:.2f} ms/step")
with torch.cuda.stream(sp.prep_stream):
    a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a0.record()
    for i in range(20):
        eng.prepare(*resident[i % 3])
    a1.record()
torch.cuda.synchronize()
print(f"prepare only: device {a0.elapsed_time(a1) / 20:.2f} ms/step (includes its host round trip)")
pr = cProfile.Profile()
pr.enable()
for i in range(10):
    sp.submit(*resident[i % 3], wait_main=False)
pr.disable()
torch.cuda.synchronize()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(35)
print(s.getvalue()[:6000])
