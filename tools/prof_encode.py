"""Small driver for ncu: runs the fused encode (link + materialise) a few times on one S1 sweep.
usage: python tools/prof_encode.py [reps] [n_points] [batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from efficient_multimodal_perception_b200 import ops, synth  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
n = int(sys.argv[2]) if len(sys.argv) > 2 else 34720
batch = int(sys.argv[3]) if len(sys.argv) > 3 else 1
dev = torch.device("cuda:0")
G = synth.GEOM_A
pts = torch.cat([synth.lidar_sweep(n, seed=1001 + b) for b in range(batch)])
xyz = pts[:, :3].contiguous().to(dev)
feats = synth.point_features(n * batch, 128, seed=1001).to(dev)
off = synth.batch_offsets([n] * batch).to(dev)
for _ in range(reps):
    out = ops.encode(feats, off, G["pc_range"], G["voxel_size"], G["grid_size"], G["split"], points=xyz)
torch.cuda.synchronize()
print("done", reps, n, batch)
