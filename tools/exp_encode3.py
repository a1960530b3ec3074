"""encode timings over point densities (1, 2, 4, 10 accumulated sweeps)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from efficient_multimodal_perception_b200 import ops, synth
dev = torch.device("cuda:0"); G = synth.GEOM_A
def timeit(fn, reps=20):
    for _ in range(4): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
for sweeps in (1, 2, 4, 10):
    pts = synth.multi_sweep(sweeps, 34720) if sweeps > 1 else synth.lidar_sweep(34720, seed=1001)
    n = pts.shape[0]
    xyz = pts[:, :3].contiguous().to(dev); feats = synth.point_features(n, 128, seed=1).to(dev); off = synth.batch_offsets([n]).to(dev)
    f = lambda: ops.encode(feats, off, G["pc_range"], G["voxel_size"], G["grid_size"], G["split"], points=xyz)
    print(f"n={n:7d} encode {timeit(f):7.1f} us", flush=True)
