"""Chunked pinned H2D copies: one stream back to back vs the chunks dealt round-robin over 2 / 3 streams (do the gaps between
consecutive copies of one stream disappear?).  python tools/pcie_probe2.py"""
import time
import torch
dev = torch.device("cuda", 0)
n = 469 * 1000 * 1000
h = torch.empty(n, dtype=torch.uint8).pin_memory()
d = torch.empty(n, dtype=torch.uint8, device=dev)
streams = [torch.cuda.Stream() for _ in range(4)]


def run(chunk, ns):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        for i, o in enumerate(range(0, n, chunk)):
            with torch.cuda.stream(streams[i % ns]):
                d[o:o + chunk].copy_(h[o:o + chunk], non_blocking=True)
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / 5


run(n, 1)
print("one copy: %.2f ms" % (run(n, 1) * 1e3))
for mb in (60, 30, 15, 8):
    print("%2d MB chunks:" % mb, "  ".join("%d stream%s %.2f ms" % (ns, "s" if ns > 1 else " ", run(mb * 1000 * 1000, ns) * 1e3) for ns in (1, 2, 3, 4)))
