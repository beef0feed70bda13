import torch, time
n = 256 << 20
h = torch.empty(n, dtype=torch.uint8).pin_memory(); d = torch.empty(n, dtype=torch.uint8, device="cuda")
h2 = torch.empty(n, dtype=torch.uint8).pin_memory(); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=10):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
a = t(lambda: d.copy_(h, non_blocking=True)); print("H2D alone GB/s", n / a / 1e9)
b = t(lambda: h2.copy_(d2, non_blocking=True)); print("D2H alone GB/s", n / b / 1e9)
def both():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
c = t(both); print("concurrent: each direction GB/s", n / c / 1e9)
def both21():
    with torch.cuda.stream(s1): d.copy_(h, non_blocking=True); d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
c = t(both21); print("2 H2D + 1 D2H: H2D GB/s", 2 * n / c / 1e9, "D2H", n / c / 1e9)
for mb in (4, 8, 16, 32, 64):
    m = mb << 20
    def chunks():
        for i in range(0, n, m): d[i:i+m].copy_(h[i:i+m], non_blocking=True)
    print("H2D chunks of", mb, "MiB GB/s", n / t(chunks, 5) / 1e9)
