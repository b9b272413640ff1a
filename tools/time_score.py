import sys, os
sys.path[:0] = [os.getcwd()]
import torch
from lidal_b200 import score, synth
seq = synth.make_sequence(25, "SK", seed=77)
sc = score.SequenceScorer("cuda")
for i in range(25):
    sc.add_frame(seq.xyz[i], synth.synthetic_probs(seq.xyz[i], 19, 300 + i), seq.sv_id[i], seq.sv2point[i])
for _ in range(3):
    sc.score_frame_device(12)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): sc.score_frame_device(12)
e1.record(); torch.cuda.synchronize()
print("cell factor", score.GRID_CELL_FACTOR, "score ms/frame", e0.elapsed_time(e1) / 10)
