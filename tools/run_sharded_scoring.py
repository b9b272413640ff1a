"""torchrun --nproc-per-node N tools/run_sharded_scoring.py : sharded LiDAL scoring on N GPUs (NCCL halo exchange + one
all_gather) must equal the single-GPU result exactly; rank 0 then runs the global selection and prints timings."""
import os, sys, time
sys.path[:0] = [os.getcwd()]
import numpy as np
import torch
import torch.distributed as dist
from lidal_b200 import pipeline, score, synth

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
n_frames, n_cls = int(os.environ.get("N_FRAMES", 64)), 19
seq = synth.make_sequence(n_frames, "NU", seed=21, step=0.4)
probs = [synth.synthetic_probs(seq.xyz[i], n_cls, 900 + i) for i in range(n_frames)]
n_points = [x.shape[0] for x in seq.xyz]
n_regions = int(seq.sv_id[-1][-1]) + 1
own = pipeline.frame_shard(n_frames, world, rank)
frames = {f: (torch.from_numpy(seq.xyz[f]).to(dev), torch.from_numpy(probs[f]).to(dev)) for f in own}
regions = {f: (seq.sv_id[f], seq.sv2point[f]) for f in own}
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
d, e, pn, c = pipeline.score_sequence_sharded(frames, n_points, n_cls, n_frames, regions,
                                              pipeline.cuda_score_frame(n_frames, device=dev), n_regions, device=dev)
torch.cuda.synchronize()
t_sharded = time.perf_counter() - t0
if rank == 0:
    sc = score.SequenceScorer(dev)
    for i in range(n_frames):
        sc.add_frame(seq.xyz[i], probs[i], seq.sv_id[i], seq.sv2point[i])
    D, E, PN, C = np.zeros(n_regions, np.float32), np.zeros(n_regions, np.float32), np.zeros(n_regions, int), np.zeros((n_regions, 3), np.float32)
    t0 = time.perf_counter()
    for i in range(n_frames):
        sv_id, sd, se, sn, scn = sc.score_frame(i)
        D[sv_id], E[sv_id], PN[sv_id], C[sv_id] = sd, se, sn, scn
    t_single = time.perf_counter() - t0
    ok = np.array_equal(D, d) and np.array_equal(E, e) and np.array_equal(PN, pn) and np.array_equal(C, c)
    flags0 = (np.random.default_rng(0).random(n_regions) < 0.01).astype(np.float64)
    t0 = time.perf_counter()
    flags = score.select_regions(flags0, d, e, pn, c, train_point_num=int(pn.sum()), device=dev)
    t_sel = time.perf_counter() - t0
    print(f"world={world} frames={n_frames} regions={n_regions} sharded==single: {ok}  sharded {t_sharded*1e3:.1f} ms "
          f"single {t_single*1e3:.1f} ms  selection {t_sel*1e3:.1f} ms  labelled={int((flags==1).sum())} pseudo={int((flags==2).sum())}")
    assert ok
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
