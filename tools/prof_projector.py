"""Stage timing of PointTriplaneProjector.forward at bs=8 (range_cam workload): where do the ms go?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import efficient_multimodal_perception_b200 as emp
from efficient_multimodal_perception_b200 import ops, synth
from torch.profiler import profile, ProfilerActivity

dev = torch.device("cuda:0")
B, n = 8, 34720
G = synth.GEOM_A
C = G["channels"]
torch.manual_seed(0)
proj = emp.PointTriplaneProjector(G["grid_size"], in_channels=5, out_channels=C, base_channels=C, split=G["split"]).eval().to(dev)
pts = [synth.lidar_sweep(n, seed=1004 + b).to(dev) for b in range(B)]
with torch.no_grad():
    cropped, gi = emp.voxelize_points(pts, G["pc_range"], G["voxel_size"])
    cam = [torch.randn(c.shape[0], 768, device=dev) for c in cropped]
    for _ in range(3):
        proj(cropped, gi, cam)
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(5):
            proj(cropped, gi, cam)
        torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
