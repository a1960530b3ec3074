"""Timings of the backward kernels next to torch-CUDA autograd of the reference op chain."""
import os, sys, torch
import torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from efficient_multimodal_perception_b200 import ops, synth
dev = torch.device("cuda:0")
LO, VS, HALF = [-25.0, -25.0, -5.0], (0.4, 0.4, 0.1), [64.0] * 3
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
def torch_chain(planes, q):
    v = torch.zeros_like(q)
    for a in range(3): v[..., a] = (q[..., a] - LO[a]) / VS[a]
    for a in range(3): v[..., a] = v[..., a] / HALF[a] - 1
    v = v[:, None]
    return (F.grid_sample(planes[0], v[..., [0, 1]], align_corners=False) + F.grid_sample(planes[1], v[..., [1, 2]], align_corners=False)
            + F.grid_sample(planes[2], v[..., [0, 2]], align_corners=False))[:, :, 0]
for name, q in (("lattice640k", synth.occ_gt_lattice().reshape(1, -1, 3)), ("uniform640k", synth.uniform_queries(640000)[None]),
                ("range_bs1", synth.range_image_points(1).reshape(1, -1, 3))):
    q = q.contiguous().to(dev)
    tri = synth.triplane_stacked(1, 32, 128, seed=1).to(dev).requires_grad_()
    g = torch.randn(1, 32, q.shape[1], device=dev)
    t_ours = timeit(lambda: ops.sample3_backward(g, q, [(128, 128)] * 3, LO, VS, HALF))
    if name == "lattice640k":
        t_grid = timeit(lambda: ops.sample3_backward(g, q, [(128, 128)] * 3, LO, VS, HALF, grid_dims=(200, 200, 16)))
        print(f"decode backward {name:12s}: lattice kernel {t_grid:8.1f} us (incl. zeroing + NHWC->NCHW of the gradient planes)")
    planes = [tri[:, 0], tri[:, 1], tri[:, 2]]
    def tb():
        out = torch_chain(planes, q)
        return torch.autograd.grad(out, [tri], g)
    t_fb = timeit(tb)
    t_f = timeit(lambda: torch_chain([p.detach() for p in planes], q))
    print(f"decode backward {name:12s}: ours {t_ours:8.1f} us   torch fwd+bwd {t_fb:8.1f} us (fwd alone {t_f:8.1f} us)")
G = synth.GEOM_A
pts = synth.lidar_sweep(34720, seed=1001)
keep, idx = ops.voxel_index(pts[:, :3].contiguous().to(dev), G["pc_range"], G["voxel_size"])
idx = idx[keep.bool()].contiguous(); n = idx.shape[0]
feats = synth.point_features(n, 128, seed=1).to(dev)
off = synth.batch_offsets([n]).to(dev)
outs = ops.encode(feats, off, [0] * 6, (1, 1, 1), G["grid_size"], G["split"], grid_ind=idx)
gs = [torch.randn_like(o) for o in outs]
print(f"encode backward (max, {n} pts): {timeit(lambda: ops.encode_backward(gs, feats, idx, off, G['grid_size'], G['split'], outs=outs)):8.1f} us")
