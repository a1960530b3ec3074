"""Small driver for ncu: the fused Mlp head on 640k queries."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from efficient_multimodal_perception_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
x = torch.randn(1, 32, 640000, generator=g).to(dev)
w1 = torch.randn(64, 32, generator=g).to(dev); w2 = torch.randn(32, 64, generator=g).to(dev); w3 = torch.randn(5, 32, generator=g).to(dev)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    ops.mlp_head(x, w1, w2, w3)
torch.cuda.synchronize()
print("done")
