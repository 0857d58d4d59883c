"""Raw pinned-memory copy bandwidth of the box (ceiling of the e2e number)."""
import time, torch
n = 64 << 20
h1 = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d1 = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(fn, reps=20):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return reps * n / (time.perf_counter() - t0) / 1e9
print("H2D 64MiB GB/s", run(lambda: d1.copy_(h1, non_blocking=True)))
print("D2H 64MiB GB/s", run(lambda: h2.copy_(d2, non_blocking=True)))
def both():
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
print("H2D+D2H concurrent, GB/s per direction", run(both))
for mib in (1, 4, 16):
    m = mib << 20
    def small():
        with torch.cuda.stream(s1): d1[:m].copy_(h1[:m], non_blocking=True)
        with torch.cuda.stream(s2): h2[:m].copy_(d2[:m], non_blocking=True)
    small(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(200): small()
    torch.cuda.synchronize(); print(f"{mib} MiB chunks both directions: {200*m/(time.perf_counter()-t0)/1e9:.1f} GB/s per direction")
