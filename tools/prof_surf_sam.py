"""Small driver for ncu: the surf_sam step (bs=8 range points through the per-query kernel + the SAM subsets through the
segment kernel), a few repetitions. usage: python tools/prof_surf_sam.py [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import efficient_multimodal_perception_b200 as emp  # noqa: E402
from efficient_multimodal_perception_b200 import ops, synth  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda:0")
B = 8
G = synth.GEOM_A
lo, vs = G["pc_range"], G["voxel_size"]
tri = synth.triplane_stacked(B, 32, 128, seed=1003).to(dev)
rp = synth.range_image_points(B, seed=1003).to(dev)
pts = [p.to(dev) for p in bench.sam_points(B, 1003)]
coords, labels, bidx = emp.sam_subsets(pts, lo)
q_cat = torch.cat(coords).contiguous()
seg_off = synth.batch_offsets([c.shape[0] for c in coords]).to(dev)
seg_b = torch.tensor(bidx, dtype=torch.int32, device=dev)
half = [64.0] * 3
for _ in range(reps):
    nhwc = ops.planes_to_channels_last([tri[:, 0], tri[:, 1], tri[:, 2]])
    a = ops.sample3(nhwc, rp.view(B, -1, 3), lo[:3], vs, half, channels_last=True)
    b = ops.sample3_segments(nhwc, q_cat, seg_off, seg_b, lo[:3], vs, half, channels_last=True)
torch.cuda.synchronize()
print("done", reps, q_cat.shape)
