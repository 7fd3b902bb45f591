"""Time the tensor-core matcher alone (orb_match_knn2_batch on resident descriptors) and check it against a torch popcount
reference on a few pairs: python tools/match_probe.py [frames] [descriptors per frame]"""
import importlib
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
V = importlib.import_module("visual-odometry-gpu_b200")
F = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
CAP = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
dev = torch.device("cuda:0")
ctx = V.Context(V.make_params(nfeatures=CAP, max_width=1241, max_height=376, max_batch=8, max_keypoints=CAP))
g = torch.Generator(device="cpu").manual_seed(7)
d = torch.randint(0, 256, (F, CAP, 32), dtype=torch.uint8, generator=g).to(dev)
n = torch.randint(CAP - 8, CAP + 1, (F,), dtype=torch.int32, generator=g).to(dev)
m = torch.zeros(F - 1, CAP, 4, dtype=torch.int32, device=dev)


def run():
    ctx.match_knn2_batch_ptr(d.data_ptr(), n.data_ptr(), F, CAP, m.data_ptr())


for _ in range(2):
    run()
ctx.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
torch.cuda.synchronize()
ts = []
for _ in range(5):
    ev[0].record(torch.cuda.current_stream())
    run()
    ctx.synchronize()
    ev[1].record(torch.cuda.current_stream())
    torch.cuda.synchronize()
    ts.append(ev[0].elapsed_time(ev[1]))
print("frames", F, "cap", CAP, "ms per batch: median %.3f min %.3f" % (sorted(ts)[2], min(ts)))
# spot check, pairs 0 and F-2
bad = 0
for p in (0, F - 2):
    nq, nt = int(n[p]), int(n[p + 1])
    q = d[p, :nq].view(torch.int32).view(nq, 8)
    t = d[p + 1, :nt].view(torch.int32).view(nt, 8)
    x = (q[:, None, :] ^ t[None, :, :])
    x = x.view(torch.uint8).view(nq, nt, 32).to(torch.int16)
    lut = torch.tensor([bin(i).count("1") for i in range(256)], dtype=torch.int16, device=dev)
    dist = lut[x.long()].sum(-1).to(torch.int32)
    key = dist * 16384 + torch.arange(nt, device=dev, dtype=torch.int32)[None, :]
    k2 = torch.topk(key, 2, dim=1, largest=False).values
    got = m[p, :nq]
    ok = (got[:, 0] == (k2[:, 0] & 16383)) & (got[:, 1] == (k2[:, 0] >> 14)) & (got[:, 2] == (k2[:, 1] & 16383)) & (got[:, 3] == (k2[:, 1] >> 14))
    bad += int((~ok).sum())
print("mismatches", bad)
ctx.close()
