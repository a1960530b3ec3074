"""Experiment: decode + occupancy head as two kernels vs the fused kernel. usage: python tools/exp_head.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from efficient_multimodal_perception_b200 import ops, synth
dev = torch.device("cuda:0")


def timeit(fn, reps=200, group=20):
    """Time `group` back-to-back launches replayed from one CUDA graph (launch overhead amortised the way
    a real step sees it); returns avg / median / min per launch in us over reps // group replays."""
    for i in range(group):
        fn(i)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for i in range(group):
            fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(group):
            fn(i)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    evs = []
    for _ in range(max(3, reps // group)):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) / group for a, b in evs)
    return sum(ts) / len(ts) * 1e3, ts[len(ts) // 2] * 1e3, ts[0] * 1e3


LO, VS, HALF = [-25.0, -25.0, -5.0], (0.4, 0.4, 0.1), [64.0] * 3
g = torch.Generator().manual_seed(5)
w1 = (torch.randn(64, 32, generator=g) / 32 ** 0.5).to(dev); w2 = (torch.randn(32, 64, generator=g) / 8).to(dev); w3 = (torch.randn(5, 32, generator=g) / 32 ** 0.5).to(dev)
for name, q, B in (("lattice640k", synth.occ_gt_lattice(), 1), ("roi", synth.roi_lattice(), 1), ("roi_x8", synth.roi_lattice(), 8)):
    dims = tuple(q.shape[:3])
    qd = q.reshape(1, -1, 3).repeat(B, 1, 1).to(dev)
    nsets = 4
    tris = [synth.triplane_stacked(B, 32, 128, seed=10 + s).to(dev) for s in range(nsets)]
    nhwc = [ops.planes_to_channels_last([t[:, 0], t[:, 1], t[:, 2]]) for t in tris]
    feats = [torch.empty(B, 32, qd.shape[1], device=dev) for _ in range(nsets)]
    def two(i):
        ops.sample3(nhwc[i % nsets], qd, LO, VS, HALF, channels_last=True, out=feats[i % nsets], grid_dims=dims)
        ops.mlp_head(feats[i % nsets], w1, w2, w3)
    def one(i):
        ops.sample3_head(nhwc[i % nsets], qd, LO, VS, HALF, w1, w2, w3, grid_dims=dims, channels_last=True)
    a = ops.mlp_head(ops.sample3(nhwc[0], qd, LO, VS, HALF, channels_last=True, grid_dims=dims), w1, w2, w3)
    b = ops.sample3_head(nhwc[0], qd, LO, VS, HALF, w1, w2, w3, grid_dims=dims, channels_last=True)
    t2, t1 = timeit(two), timeit(one)
    print(f"{name:12s} B={B} Q={qd.shape[1]}: two kernels {t2[0]:7.1f} us   fused {t1[0]:7.1f} us   equal: {torch.equal(a, b)}  max diff {float((a-b).abs().max()):.3e}")
