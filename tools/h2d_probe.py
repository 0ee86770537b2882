"""Host link probe: pinned -> HBM upload rate of a Houston-sized raster (386 MB fp32) through 1 / 2 / 4 streams."""
import torch, time
dev = "cuda:0"
H, W, C = 349, 1905, 145
x = torch.rand(H, W, C).pin_memory()
y = torch.empty(H, W, C, device=dev)
for ns in (1, 2, 4, 8):
    streams = [torch.cuda.Stream() for _ in range(ns)]
    rows = [(H * k) // ns for k in range(ns + 1)]
    def go():
        for s, a, b in zip(streams, rows[:-1], rows[1:]):
            with torch.cuda.stream(s):
                y[a:b].copy_(x[a:b], non_blocking=True)
    for _ in range(2):
        go()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        go()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    print(f"{ns} stream(s): {dt * 1e3:.2f} ms per 386 MB = {x.numel() * 4 / dt / 1e9:.1f} GB/s")
z = torch.empty(H, W, 16)
zp = z.pin_memory()
d = torch.rand(H, W, 16, device=dev)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): zp.copy_(d, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 5
print(f"D2H 42 MB: {dt * 1e3:.2f} ms = {zp.numel() * 4 / dt / 1e9:.1f} GB/s")
