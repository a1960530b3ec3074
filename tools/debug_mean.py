import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from efficient_multimodal_perception_b200 import ops, synth
from test_gpu_parity import _rand_encode_case, cu
grid, split, C = [32, 32, 16], [4, 4, 4], 32
inds, feats = _rand_encode_case((6000, 5000), grid, C, seed=17, hot=True)
off = cu(synth.batch_offsets([6000, 5000]))
out = ops.encode(cu(feats), off, [0]*6, (1,1,1), grid, split, grid_ind=cu(torch.cat(inds)), reduce="mean")
xy, yz, xz, cnt = ops.encode(cu(feats), off, [0]*6, (1,1,1), grid, split, grid_ind=cu(torch.cat(inds)), reduce="sum", want_counts=True)
n0, n1 = xy.numel()//C, yz.numel()//C
print("cnt max", int(cnt.max()), "sum", int(cnt.sum()))
for name, a, b, c in (("xy", xy, out[0], cnt[:n0]), ("yz", yz, out[1], cnt[n0:n0+n1]), ("xz", xz, out[2], cnt[n0+n1:])):
    a = a.clone(); ops.finalize_mean(a, c.contiguous(), C)
    d = (a-b).abs()
    print(name, "maxdiff", float(d.max()), "max|b|", float(b.abs().max()), "nan", bool(torch.isnan(a).any()), bool(torch.isnan(b).any()))
