"""torchrun --nproc-per-node N tools/ddp_train_check.py : MinkUNet DDP training steps through lidal_b200.compat (config 4:
batch 2 SK-shaped scans per GPU, NCCL gradient all-reduce).  Checks that the ranks hold identical parameters after the
steps (gradients were all-reduced) and prints ms/step (max over ranks)."""
import os, sys
sys.path[:0] = [os.getcwd()]
import torch, torch.distributed as dist
import lidal_b200.compat as ts
from lidal_b200 import synth
from lidal_b200.network import MinkUNet, seeded_state_dict
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
c, f, _ = synth.scan_batch(seed=5 + rank, kind="SK", batch=2)          # every rank trains on its own scans
coords, feats = torch.from_numpy(c).to(dev), torch.from_numpy(f).to(dev)
labels = torch.randint(0, 19, (coords.shape[0],), device=dev)
model = MinkUNet(19, ts)
model.load_state_dict(seeded_state_dict(model.state_dict()))
model = torch.nn.parallel.DistributedDataParallel(model.to(dev).train(), device_ids=[local], output_device=local)
opt = torch.optim.Adam(model.parameters())
def step():
    opt.zero_grad()
    logits, _ = model(ts.SparseTensor(feats, coords))
    loss = torch.nn.functional.cross_entropy(logits, labels, ignore_index=255)
    loss.backward(); opt.step()
    return loss
for _ in range(3): step()
dist.barrier(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): loss = step()
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / 10], device=dev)
dist.all_reduce(ms, op=dist.ReduceOp.MAX)
# parameters must be bit-identical across ranks after all-reduced updates
flat = torch.cat([p.detach().flatten() for p in model.parameters()])
ref = flat.clone(); dist.broadcast(ref, 0)
same = torch.tensor([float(torch.equal(flat, ref))], device=dev); dist.all_reduce(same, op=dist.ReduceOp.MIN)
if rank == 0:
    print(f"DDP world={world}: {float(ms):.1f} ms/step (fwd+bwd+allreduce+Adam, 2 SK scans/GPU), params identical across ranks: {bool(same.item())}, loss {float(loss):.4f}")
    assert bool(same.item())
dist.destroy_process_group()
