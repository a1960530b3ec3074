"""Small driver for ncu: runs the lift kernel a few times at the PointTriplane config (6 x [768,16,32], 2 sweeps)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from efficient_multimodal_perception_b200 import ops, synth  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
dev = torch.device("cuda:0")
B, n = 2, 34720
rig = synth.camera_rig(1004)
metas = [dict(img_shape=rig.img_shape, lidar2image=rig.lidar2image.numpy(), imgs_aug=rig.imgs_aug) for _ in range(B)]
pts = torch.cat([synth.lidar_sweep(n, seed=1004 + b)[:, :3] for b in range(B)]).contiguous().to(dev)
off = synth.batch_offsets([n] * B).to(dev)
feats = torch.randn(B, 6, 768, 16, 32, generator=torch.Generator().manual_seed(1004)).to(dev)
cams = ops.pack_cameras(metas, dev)
nhwc = ops.features_to_channels_last(feats)
for _ in range(reps):
    out = ops.lift_cam(pts, off, nhwc, cams, rig.img_shape[::-1], channels_last=True)
torch.cuda.synchronize()
print("done", reps)
