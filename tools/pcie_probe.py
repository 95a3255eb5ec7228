"""PCIe probe: H2D, D2H and both at once from pinned memory (GB/s).  Context for bench.py's e2e number."""
import time, torch
n = 256 << 20
h1 = torch.empty(n, dtype=torch.uint8).pin_memory(); h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
d1 = torch.empty(n, dtype=torch.uint8, device="cuda"); d2 = torch.empty(n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(fn, reps=5):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / reps
def h2d():
    with torch.cuda.stream(s1): d1.copy_(h1, non_blocking=True)
def d2h():
    with torch.cuda.stream(s2): h2.copy_(d2, non_blocking=True)
def both(): h2d(); d2h()
a, b, c = run(h2d), run(d2h), run(both)
print(f"H2D {n/a/1e9:.1f} GB/s  D2H {n/b/1e9:.1f} GB/s  both at once: {2*n/c/1e9:.1f} GB/s total ({c*1e3:.2f} ms for {n>>20} MiB each way)")
def chunks(k):
    m = n // k
    def f():
        for i in range(k):
            with torch.cuda.stream(s1): d1[i*m:(i+1)*m].copy_(h1[i*m:(i+1)*m], non_blocking=True)
            with torch.cuda.stream(s2): h2[i*m:(i+1)*m].copy_(d2[i*m:(i+1)*m], non_blocking=True)
    return f
for k in (8, 64, 512):
    c = run(chunks(k)); print(f"both, {k} chunks each way: {2*n/c/1e9:.1f} GB/s total")
