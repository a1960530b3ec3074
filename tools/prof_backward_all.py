"""Small driver for ncu: every backward kernel once or twice on config-sized inputs — decode backward (per-query on
640k uniform queries, lattice on the 640k grid), segment backward (48 SAM subsets), encode max-backward (one sweep,
geometry A), lift backward (6 cameras, C = 128). usage: python tools/prof_backward_all.py [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import efficient_multimodal_perception_b200 as emp  # noqa: E402
from efficient_multimodal_perception_b200 import ops, synth  # noqa: E402

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
dev = torch.device("cuda:0")
LO, VS, HALF = [-25.0, -25.0, -5.0], (0.4, 0.4, 0.1), [64.0] * 3
G = synth.GEOM_A
# decode backward
qu = synth.uniform_queries(640000)[None].contiguous().to(dev)
ql = synth.occ_gt_lattice().reshape(1, -1, 3).contiguous().to(dev)
g = torch.randn(1, 32, 640000, device=dev)
# segments (configs[2])
B = 8
pts = [p.to(dev) for p in bench.sam_points(B, 1003)]
coords, labels, bidx = emp.sam_subsets(pts, G["pc_range"])
q_cat = torch.cat(coords).contiguous()
seg_off = synth.batch_offsets([c.shape[0] for c in coords]).to(dev)
seg_b = torch.tensor(bidx, dtype=torch.int32, device=dev)
gs = torch.randn(q_cat.shape[0], 32, device=dev)
# encode backward
raw = synth.lidar_sweep(34720, seed=1001)
xyz = raw[:, :3].contiguous().to(dev)
off1 = synth.batch_offsets([raw.shape[0]]).to(dev)
cropped, ind, offs = ops.voxelize(xyz, off1, G["pc_range"], G["voxel_size"], G["grid_size"])[:3]
feats = synth.point_features(cropped.shape[0], 128, seed=5).to(dev)
outs = ops.encode(feats, offs, G["pc_range"], G["voxel_size"], G["grid_size"], G["split"], grid_ind=ind)
grads = [torch.randn_like(o) for o in outs]
# lift backward
rig = synth.camera_rig(11)
metas = [dict(img_shape=rig.img_shape, lidar2image=rig.lidar2image.numpy(), imgs_aug=rig.imgs_aug)]
cams = ops.pack_cameras(metas, dev)
lp = synth.lidar_sweep(34720, seed=13)[:, :5].contiguous().to(dev)
gl = torch.randn(lp.shape[0], 128, device=dev)
for _ in range(reps):
    ops.sample3_backward(g, qu, [(128, 128)] * 3, LO, VS, HALF)
    ops.sample3_backward(g, ql, [(128, 128)] * 3, LO, VS, HALF, grid_dims=(200, 200, 16))
    ops.sample3_segments_backward(gs, q_cat, seg_off, seg_b, B, [(128, 128)] * 3, LO, VS, HALF)
    ops.encode_backward(grads, feats, ind, offs, G["grid_size"], G["split"], outs=outs)
    ops.lift_cam_backward(gl, lp, off1, (1, 6, 128, 16, 32), cams, (256, 512))
torch.cuda.synchronize()
print("done", reps)
