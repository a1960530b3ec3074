"""Small driver for ncu: decode backward on the 640k lattice, per-query kernel and lattice kernel."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from efficient_multimodal_perception_b200 import ops, synth
dev = torch.device("cuda:0")
LO, VS, HALF = [-25.0, -25.0, -5.0], (0.4, 0.4, 0.1), [64.0] * 3
q = synth.occ_gt_lattice().reshape(1, -1, 3).contiguous().to(dev)
g = torch.randn(1, 32, q.shape[1], device=dev)
for _ in range(3):
    ops.sample3_backward(g, q, [(128, 128)] * 3, LO, VS, HALF)
    ops.sample3_backward(g, q, [(128, 128)] * 3, LO, VS, HALF, grid_dims=(200, 200, 16))
roi = synth.roi_lattice().reshape(1, -1, 3).contiguous().to(dev)
g2 = torch.randn(1, 32, roi.shape[1], device=dev)
for _ in range(3):
    ops.sample3_backward(g2, roi, [(128, 128)] * 3, LO, VS, HALF)
    ops.sample3_backward(g2, roi, [(128, 128)] * 3, LO, VS, HALF, grid_dims=(99, 99, 16))
torch.cuda.synchronize()
print("done")
