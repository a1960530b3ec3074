"""How fast can 512-byte rows be gathered from a 179 MB array in random order? (torch index_select as a neutral probe)"""
import torch
dev = torch.device("cuda:0")
def timeit(fn, reps=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3
for n in (34720, 350000, 2000000):
    feats = torch.randn(n, 128, device=dev)
    m = 3 * n
    idx = torch.randint(0, n, (m,), device=dev)
    out = torch.empty(m, 128, device=dev)
    t = timeit(lambda: torch.index_select(feats, 0, idx, out=out))
    print(f"rows={n} gathers={m}: random {t:8.1f} us  read {m*512/t/1e3:7.0f} GB/s (+ same written)")
    idx2 = torch.sort(idx).values
    t = timeit(lambda: torch.index_select(feats, 0, idx2, out=out))
    print(f"rows={n} gathers={m}: sorted {t:8.1f} us  read {m*512/t/1e3:7.0f} GB/s")
