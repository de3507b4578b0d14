"""Raw pinned-memory copy rates of the box (reference point for bench.py's e2e figure)."""
import torch, time
dev = torch.device("cuda:0")
for mb in (16, 128, 1024):
    n = mb << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    h2 = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    d2 = torch.empty(n, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
    def run(fn, reps=8):
        fn(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps): fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) / reps
    def d2h():
        with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
    def h2d():
        with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
    def both():
        d2h(); h2d()
    print("%5d MiB  d2h %.1f GB/s  h2d %.1f GB/s  both: %.1f GB/s each way" % (
        mb, n / run(d2h) / 1e9, n / run(h2d) / 1e9, n / run(both) / 1e9))
