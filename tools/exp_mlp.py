"""Timing of the fused tensor-core Mlp head vs the reference head (three 1x1x1 Conv3d) on the decode output."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from efficient_multimodal_perception_b200 import ops
dev = torch.device("cuda:0")
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
C = 32
for shape in ((1, C, 200, 200, 16), (1, C, 99, 99, 16), (8, C, 99, 99, 16)):
    x = torch.randn(*shape, device=dev)
    head = torch.nn.Sequential(torch.nn.Conv3d(C, 2 * C, 1, bias=False), torch.nn.ReLU(inplace=True),
                               torch.nn.Conv3d(2 * C, C, 1, bias=False), torch.nn.ReLU(inplace=True),
                               torch.nn.Conv3d(C, 5, 1, bias=False)).to(dev)
    w1, w2, w3 = head[0].weight.detach(), head[2].weight.detach(), head[4].weight.detach()
    with torch.no_grad():
        t_ref = timeit(lambda: head(x))
        t_cl = timeit(lambda: head.to(memory_format=torch.channels_last_3d)(x.contiguous(memory_format=torch.channels_last_3d)))
    t_ours = timeit(lambda: ops.mlp_head(x, w1, w2, w3))
    Q = x.numel() // C
    print(f"{tuple(shape)}: ours {t_ours:8.1f} us ({Q / t_ours / 1e3:.2f} G queries/s, {Q * (4 * C + 20) / t_ours / 1e3:.0f} GB/s)   "
          f"torch Conv3d head {t_ref:8.1f} us   (channels_last_3d incl. conversion {t_cl:8.1f} us)")
