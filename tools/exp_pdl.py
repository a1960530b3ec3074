"""Experiment: what the NCHW -> channels-last conversion costs in front of the decode kernels, with and without the
programmatic dependent launch, per library build. usage: python tools/exp_pdl.py [path/to/libtriplane_variant.so]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import efficient_multimodal_perception_b200._lib as L  # noqa: E402

if len(sys.argv) > 1:
    L.LIB_PATH = os.path.abspath(sys.argv[1])
from efficient_multimodal_perception_b200 import ops, synth  # noqa: E402

dev = torch.device("cuda:0")
LO, VS, HALF = [-25.0, -25.0, -5.0], (0.4, 0.4, 0.1), [64.0] * 3


def timeit(fn, nsets, reps=30):
    for i in range(nsets):
        fn(i)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for i in range(nsets):
            fn(i)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(nsets):
            fn(i)
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    evs = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) / nsets for a, b in evs)
    return ts[len(ts) // 2] * 1e3


def case(name, B, q, dims):
    Q = q.shape[1]
    per = B * (Q * 140 + 2 * 4 * 32 * 3 * 128 * 128)
    nsets = max(4, -(-3 * 126 * 2**20 // per))
    tris = [synth.triplane_stacked(B, 32, 128, seed=1002 + s).to(dev) for s in range(nsets)]
    qs = [q.to(dev).clone() for _ in range(nsets)]
    outs = [torch.empty(B, 32, Q, device=dev) for _ in range(nsets)]
    nhwc = [ops.planes_to_channels_last([t[:, 0], t[:, 1], t[:, 2]]) for t in tris]

    def conv(i):
        t = tris[i % nsets]
        ops.planes_to_channels_last([t[:, 0], t[:, 1], t[:, 2]])

    def dec(i):
        ops.sample3(nhwc[i % nsets], qs[i % nsets], LO, VS, HALF, channels_last=True, out=outs[i % nsets], grid_dims=dims)

    def both_pdl(i):
        ops.sample3(tris[i % nsets], qs[i % nsets], LO, VS, HALF, out=outs[i % nsets], grid_dims=dims)

    def both_plain(i):
        t = tris[i % nsets]
        n = ops.planes_to_channels_last([t[:, 0], t[:, 1], t[:, 2]])
        ops.sample3(n, qs[i % nsets], LO, VS, HALF, channels_last=True, out=outs[i % nsets], grid_dims=dims)

    r = {k: timeit(f, nsets) for k, f in (("conv", conv), ("decode", dec), ("conv+decode pdl", both_pdl), ("conv+decode plain", both_plain))}
    print(f"{name:14s} B={B} Q={Q}: " + "  ".join(f"{k} {v:6.2f} us" for k, v in r.items()), flush=True)


print("lib:", L.LIB_PATH)
lat = synth.occ_gt_lattice()
case("lattice640k", 1, lat.reshape(1, -1, 3), tuple(lat.shape[:3]))
case("uniform640k", 1, synth.uniform_queries(640000)[None], None)
roi = synth.roi_lattice()
case("roi", 1, roi.reshape(1, -1, 3), tuple(roi.shape[:3]))
case("range_bs8", 8, synth.range_image_points(8).reshape(8, -1, 3), None)
