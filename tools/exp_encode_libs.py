"""Experiment: encode step time (10-sweep x 2 samples, single sweep x 8 samples) per library build.
usage: python tools/exp_encode_libs.py lib1.so [lib2.so ...]   (each library in its own process)"""
import os
import subprocess
import sys

if len(sys.argv) > 2 or (len(sys.argv) == 2 and not sys.argv[1].startswith("--one=")):
    for lib in sys.argv[1:]:
        subprocess.run([sys.executable, __file__, "--one=" + lib], check=False)
    sys.exit(0)

import torch  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import efficient_multimodal_perception_b200._lib as L  # noqa: E402

lib = sys.argv[1][len("--one="):]
L.LIB_PATH = os.path.abspath(lib)
from efficient_multimodal_perception_b200 import ops, synth  # noqa: E402

dev = torch.device("cuda:0")
G = synth.GEOM_A


def run(clouds, reps=12):
    xyz = torch.cat(clouds)[:, :3].contiguous().to(dev)
    n = [c.shape[0] for c in clouds]
    feats = synth.point_features(sum(n), 128, seed=1001).to(dev)
    off = synth.batch_offsets(n).to(dev)
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ops.encode(feats, off, G["pc_range"], G["voxel_size"], G["grid_size"], G["split"], points=xyz)
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts = sorted(ts[3:])
    return ts[len(ts) // 2] * 1e3


dense = run([synth.multi_sweep(10, 35000, seed=1005 + b) for b in range(2)])
single = run([synth.lidar_sweep(34720, seed=1001 + b) for b in range(8)])
one = run([synth.lidar_sweep(34720, seed=1001)])
print(f"{os.path.basename(lib):28s} 10-sweep x2: {dense:7.1f} us   1-sweep x8: {single:7.1f} us   1-sweep x1: {one:6.1f} us", flush=True)
