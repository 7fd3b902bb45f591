"""PCIe copy bandwidth of the box: pinned H2D alone, D2H alone, both at once (what bounds bench.py's e2e figure)."""
import time
import torch
dev = torch.device("cuda", 0)
n = 469 * 1000 * 1000
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device=dev)
h2 = torch.empty(88 * 1000 * 1000, dtype=torch.uint8).pin_memory()
d2 = torch.empty(88 * 1000 * 1000, dtype=torch.uint8, device=dev)
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def run(up, down, chunk=None):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        if up:
            with torch.cuda.stream(s1):
                if chunk:
                    for o in range(0, n, chunk):
                        d[o:o + chunk].copy_(h[o:o + chunk], non_blocking=True)
                else:
                    d.copy_(h, non_blocking=True)
        if down:
            with torch.cuda.stream(s2):
                h2.copy_(d2, non_blocking=True)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 5
    return dt
for _ in range(2):
    run(True, True)
t = run(True, False); print("H2D alone      %.1f GB/s" % (n / t / 1e9))
t = run(True, False, 60 * 1000 * 1000); print("H2D 60 MB chunks %.1f GB/s" % (n / t / 1e9))
t = run(False, True); print("D2H alone      %.1f GB/s" % (h2.numel() / t / 1e9))
t = run(True, True); print("both: H2D %.1f GB/s + D2H %.1f GB/s in the same %.1f ms" % (n / t / 1e9, h2.numel() / t / 1e9, t * 1e3))
