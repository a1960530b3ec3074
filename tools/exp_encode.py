import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from efficient_multimodal_perception_b200 import ops, synth
dev = torch.device("cuda:0")
G = synth.GEOM_A
def timeit(fn, reps=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
big = torch.empty(430 * 1024 * 1024 // 4, device=dev)
print("torch zero_ 430MB us", timeit(lambda: big.zero_()))
src = torch.empty_like(big)
print("torch copy_ 430MB us", timeit(lambda: big.copy_(src)))
for n, C in ((0, 128), (34720, 128), (34720, 32), (350000, 128)):
    pts = synth.lidar_sweep(max(n, 1), seed=1001)[:n] if n <= 34720 else synth.multi_sweep(10, 35000)
    n = pts.shape[0]
    xyz = pts[:, :3].contiguous().to(dev)
    feats = synth.point_features(n, C, seed=1).to(dev)
    off = synth.batch_offsets([n]).to(dev)
    f = lambda: ops.encode(feats, off, G["pc_range"], G["voxel_size"], G["grid_size"], G["split"], points=xyz)
    print(f"encode n={n} C={C} us", timeit(f))
    if n:
        shuf = torch.randperm(n, device=dev)
        xyz2, feats2 = xyz[shuf].contiguous(), feats[shuf].contiguous()
        f2 = lambda: ops.encode(feats2, off, G["pc_range"], G["voxel_size"], G["grid_size"], G["split"], points=xyz2)
        print(f"  shuffled points us", timeit(f2))
        f3 = lambda: ops.encode(feats, off, G["pc_range"], G["voxel_size"], G["grid_size"], G["split"], points=xyz, planes=(True, False, False))
        print(f"  xy plane only us", timeit(f3))
