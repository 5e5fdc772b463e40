import torch, time
n = 10_240_000
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
s = [torch.cuda.Stream() for _ in range(4)]
def run(k, reps=200):
    parts = [(i * n // k, (i + 1) * n // k) for i in range(k)]
    for _ in range(10):
        for i, (a, b) in enumerate(parts):
            with torch.cuda.stream(s[i]): d[a:b].copy_(h[a:b], non_blocking=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps):
        for i, (a, b) in enumerate(parts):
            with torch.cuda.stream(s[i]): d[a:b].copy_(h[a:b], non_blocking=True)
        torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / reps
    print(f"{k} concurrent H2D streams: {dt*1e6:.1f} us, {n/dt/1e9:.1f} GB/s")
for k in (1, 2, 4): run(k)
# D2H concurrently with H2D
hk = torch.empty(480_000, dtype=torch.uint8).pin_memory(); dk = torch.empty(480_000, dtype=torch.uint8, device="cuda")
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(200):
    with torch.cuda.stream(s[0]): d.copy_(h, non_blocking=True)
    with torch.cuda.stream(s[1]): hk.copy_(dk, non_blocking=True)
    torch.cuda.synchronize()
print(f"H2D 10.24 MB + concurrent D2H 0.48 MB: {(time.perf_counter()-t0)/200*1e6:.1f} us")
