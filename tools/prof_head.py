"""Small driver for ncu: decode + head in one kernel on the 640k lattice (or the roi lattice x8).
usage: python tools/prof_head.py [lattice640k|roi_x8] [reps]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from efficient_multimodal_perception_b200 import ops, synth
kind = sys.argv[1] if len(sys.argv) > 1 else "lattice640k"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device("cuda:0")
LO, VS, HALF = [-25.0, -25.0, -5.0], (0.4, 0.4, 0.1), [64.0] * 3
q, B = (synth.occ_gt_lattice(), 1) if kind == "lattice640k" else (synth.roi_lattice(), 8)
dims = tuple(q.shape[:3])
qd = q.reshape(1, -1, 3).repeat(B, 1, 1).to(dev)
g = torch.Generator().manual_seed(5)
w1 = (torch.randn(64, 32, generator=g) / 32 ** 0.5).to(dev); w2 = (torch.randn(32, 64, generator=g) / 8).to(dev); w3 = (torch.randn(5, 32, generator=g) / 32 ** 0.5).to(dev)
tri = synth.triplane_stacked(B, 32, 128, seed=11).to(dev)
nhwc = ops.planes_to_channels_last([tri[:, 0], tri[:, 1], tri[:, 2]])
for _ in range(reps):
    ops.sample3_head(nhwc, qd, LO, VS, HALF, w1, w2, w3, grid_dims=dims, channels_last=True)
torch.cuda.synchronize()
print("done", kind, reps)
