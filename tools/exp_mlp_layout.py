"""Debug: which (channel, query) element does every logit position see? Identity-like weights make the head a selector.
usage: python tools/exp_mlp_layout.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from efficient_multimodal_perception_b200 import ops
dev = torch.device("cuda:0")
Q = int(sys.argv[1]) if len(sys.argv) > 1 else 256
w1 = torch.zeros(64, 32); w1[:32] = torch.eye(32)
w2 = torch.zeros(32, 64); w2[:, :32] = torch.eye(32)
q = torch.arange(Q)
xc = (torch.arange(32).float() + 1).view(1, 32, 1).expand(1, 32, Q).contiguous()
xq = ((q % 128).float() + 1).view(1, 1, Q).expand(1, 32, Q).contiguous()
for base in (0, 27):
    w3 = torch.zeros(5, 32)
    for cl in range(5):
        w3[cl, base + cl] = 1
    oc = ops.mlp_head(xc.to(dev), w1.to(dev), w2.to(dev), w3.to(dev)).cpu()[0]
    oq = ops.mlp_head(xq.to(dev), w1.to(dev), w2.to(dev), w3.to(dev)).cpu()[0]
    print(f"== selecting channels {base}..{base+4} (expected channel ids {base+1}..{base+5}, row ids 1..128 repeating)")
    for cl in range(5):
        print(f" class {cl}: channel seen at q=0..15: {oc[cl, :16].int().tolist()}  distinct over all q: {sorted(set(oc[cl].int().tolist()))[:12]}")
    print("  row seen at q=0..39:", oq[0, :40].int().tolist())
    print("  row seen at q=128..167:", oq[0, 128:168].int().tolist())
