"""Pinned host -> device bandwidth: one copy stream against two and four concurrent ones."""
import time, torch
n = 1 << 30
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device="cuda")
def run(k, reps=5):
    streams = [torch.cuda.Stream() for _ in range(k)]
    part = n // k
    torch.cuda.synchronize()
    best = 0.0
    for _ in range(reps):
        t0 = time.perf_counter()
        for i, s in enumerate(streams):
            with torch.cuda.stream(s):
                d[i * part:(i + 1) * part].copy_(h[i * part:(i + 1) * part], non_blocking=True)
        torch.cuda.synchronize()
        best = max(best, n / (time.perf_counter() - t0) / 1e9)
    return best
for k in (1, 2, 4):
    print(k, "stream(s): %.1f GB/s" % run(k))
