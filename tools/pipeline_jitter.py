"""Is the occasional 2x step time of StreamPipeline a property of the side stream (hardware-queue aliasing with the main
stream) or of the caching allocator?  Builds a fresh StreamPipeline per trial and prints ms/step, cudaMalloc deltas and the
stream handle."""
import os, sys
sys.path[:0] = [os.getcwd()]
import torch
import bench
from lidal_b200.engine import StreamPipeline
dev = torch.device("cuda:0")
run, eng = bench.build_runner("spvcnn", "engine", dev)
batches = bench.make_batches(0)
resident = [(torch.from_numpy(c).to(dev), torch.from_numpy(f).to(dev)) for c, f, _ in batches]
trials = int(sys.argv[1]) if len(sys.argv) > 1 else 12
for trial in range(trials):
    sp = StreamPipeline(eng)
    for i in range(5):
        sp.submit(*resident[i % 3], wait_main=False)
    torch.cuda.synchronize()
    st0 = torch.cuda.memory_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        sp.submit(*resident[i % 3], wait_main=False)
    e1.record()
    torch.cuda.synchronize()
    st1 = torch.cuda.memory_stats()
    print(f"trial {trial:2d}: {e0.elapsed_time(e1) / 20:7.3f} ms/step  cudaMalloc +{st1['num_device_alloc'] - st0['num_device_alloc']} "
          f"cudaFree +{st1['num_device_free'] - st0['num_device_free']} retries +{st1['num_alloc_retries'] - st0['num_alloc_retries']} "
          f"reserved {st1['reserved_bytes.all.current'] >> 20} MiB  prep stream {sp.prep_stream.cuda_stream:#x}", flush=True)
