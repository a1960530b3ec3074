import torch, time
d = torch.empty(81920000//4, device="cuda"); h = torch.empty(81920000//4).pin_memory()
hq = torch.empty(13971456//4).pin_memory(); dq = torch.empty(13971456//4, device="cuda")
def t(fn, n=20):
    fn(); torch.cuda.synchronize(); t0=time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter()-t0)/n*1e3
a = t(lambda: h.copy_(d, non_blocking=True)); print("D2H 82MB ms", a, 81.92/a, "GB/s")
b = t(lambda: dq.copy_(hq, non_blocking=True)); print("H2D 14MB ms", b, 13.97/b, "GB/s")
s2 = torch.cuda.Stream()
def both():
    with torch.cuda.stream(s2): dq.copy_(hq, non_blocking=True)
    h.copy_(d, non_blocking=True)
print("both concurrently ms", t(both))
# strided D2H: [32, 160000] blocks into [32, 640000]
d2 = torch.empty(32, 160000, device="cuda"); h2 = torch.empty(32, 640000).pin_memory()
c = t(lambda: h2[:, :160000].copy_(d2, non_blocking=True)); print("D2H 2D 20MB ms", c, 20.48/c, "GB/s")
