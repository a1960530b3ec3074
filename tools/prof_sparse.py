"""Small driver for ncu: tp_projector_sparse_f32 at bs=8 (range_cam workload), a few repetitions."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import efficient_multimodal_perception_b200 as emp  # noqa: E402
from efficient_multimodal_perception_b200 import ops, synth  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device("cuda:0")
G = synth.GEOM_A
C = G["channels"]
pts = [synth.lidar_sweep(34720, seed=1004 + b).to(dev) for b in range(B)]
cropped, gi = emp.voxelize_points(pts, G["pc_range"], G["voxel_size"])
n = sum(c.shape[0] for c in cropped)
feats = synth.point_features(n, C, seed=3).to(dev)
off = synth.batch_offsets([g.shape[0] for g in gi]).to(dev)
cat = torch.cat(gi)
gen = torch.Generator().manual_seed(1)
ws = [(torch.randn(C, k * C, generator=gen) / (k * C) ** 0.5).to(dev) for k in (20, 25, 25)]
bs = [torch.randn(C, generator=gen).to(dev) for _ in range(3)]
for _ in range(reps):
    h = ops.projector_sparse(feats, off, [0] * 6, (1, 1, 1), G["grid_size"], G["split"], ws, bs, grid_ind=cat)
torch.cuda.synchronize()
print("done", reps, n)
