"""Small driver for ncu: runs the decode gather kernel a few times on one query set.
usage: python tools/prof_decode.py {uniform640k|lattice640k|roi} [reps] [grid]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from efficient_multimodal_perception_b200 import ops, synth  # noqa: E402

kind = sys.argv[1] if len(sys.argv) > 1 else "uniform640k"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
use_grid = len(sys.argv) > 3 and sys.argv[3] == "grid"
dims = {"lattice640k": (200, 200, 16), "roi": (99, 99, 16)}.get(kind) if use_grid else None
dev = torch.device("cuda:0")
q = bench.decode_queries(kind).to(dev)
nsets = 4
tris = [synth.triplane_stacked(1, 32, 128, seed=1002 + s).to(dev) for s in range(nsets)]
nhwc = [ops.planes_to_channels_last([t[:, 0], t[:, 1], t[:, 2]]) for t in tris]
outs = [torch.empty(1, 32, q.shape[1], device=dev) for _ in range(nsets)]
for i in range(reps):
    ops.sample3(nhwc[i % nsets], q, bench.OCC_LO, bench.OCC_VS, bench.OCC_HALF, channels_last=True, out=outs[i % nsets],
                grid_dims=dims)
torch.cuda.synchronize()
print("done", kind, reps)
