"""Small driver for ncu: the fused encode on 10-sweep samples (~350k points each, BASELINE configs[4]).
usage: python tools/prof_encode_dense.py [reps] [batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from efficient_multimodal_perception_b200 import ops, synth  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
batch = int(sys.argv[2]) if len(sys.argv) > 2 else 2
dev = torch.device("cuda:0")
G = synth.GEOM_A
clouds = [synth.multi_sweep(10, 35000, seed=1005 + b) for b in range(batch)]
xyz = torch.cat(clouds)[:, :3].contiguous().to(dev)
n = [c.shape[0] for c in clouds]
feats = synth.point_features(sum(n), 128, seed=1001).to(dev)
off = synth.batch_offsets(n).to(dev)
evs = []
for _ in range(reps):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    out = ops.encode(feats, off, G["pc_range"], G["voxel_size"], G["grid_size"], G["split"], points=xyz)
    b.record()
    evs.append((a, b))
torch.cuda.synchronize()
print("done", reps, batch, "ms per call:", [round(a.elapsed_time(b), 3) for a, b in evs])
