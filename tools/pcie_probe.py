import torch, time
dev = torch.device("cuda:0")
n = 1 << 30
d = torch.empty(n, dtype=torch.uint8, device=dev)
h = torch.empty(n, dtype=torch.uint8).pin_memory()
h2 = torch.empty(n // 4, dtype=torch.uint8).pin_memory()
d2 = torch.empty(n // 4, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps
dt = t(lambda: h.copy_(d, non_blocking=True)); print(f"D2H 1 GiB: {n / dt / 1e9:.1f} GB/s")
dt = t(lambda: d.copy_(h, non_blocking=True)); print(f"H2D 1 GiB: {n / dt / 1e9:.1f} GB/s")
def both():
    with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
dt = t(both); print(f"D2H 1 GiB with concurrent H2D 256 MiB: D2H-equivalent {n / dt / 1e9:.1f} GB/s")
for mb in (4, 16, 64):
    m = mb << 20
    def chunks():
        for o in range(0, n, m): h[o:o + m].copy_(d[o:o + m], non_blocking=True)
    dt = t(chunks, 3); print(f"D2H in {mb} MiB chunks: {n / dt / 1e9:.1f} GB/s")
