"""Are there host-side launch stalls on this box that have nothing to do with our code?  Launches a trivial kernel in a loop
for ~25 s and prints every iteration whose host time exceeds 5 ms, with its time stamp (periodicity = an external poller)."""
import time, sys
import torch
x = torch.zeros(1024, device="cuda")
torch.cuda.synchronize()
t_start = time.perf_counter()
n = 0
stalls = []
while time.perf_counter() - t_start < float(sys.argv[1]) if len(sys.argv) > 1 else 25.0:
    t0 = time.perf_counter()
    x.add_(1)
    if n % 64 == 63:
        torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    if dt > 0.005:
        stalls.append((t0 - t_start, dt))
    n += 1
print(f"{n} launches, {len(stalls)} host stalls > 5 ms")
for ts, dt in stalls:
    print(f"  t = {ts:7.3f} s   {dt * 1e3:7.1f} ms")
