"""Batch tracker alone on synthetic KITTI-like frames: python tools/lk_probe.py [frames] [shift]   (ncu target)
shift = 0: frame t+1 is frame t moved by (1, 2) pixels (tracks converge in a few iterations); 1: the bench's frames (17, 113)."""
import importlib
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
V = importlib.import_module("visual-odometry-gpu_b200")
F = int(sys.argv[1]) if len(sys.argv) > 1 else 65
big = int(sys.argv[2]) if len(sys.argv) > 2 else 1
W, H, cap = 1241, 376, 2000
PITCH = (W + 1 + 15) // 16 * 16
if big:
    frames = V.synth_frames(F, W, H, pitch=PITCH)
else:
    base = V.synth_frames(1, W, H, pitch=PITCH)[0]
    frames = np.stack([np.roll(base, (t, 2 * t), (0, 1)) for t in range(F)])
dev = torch.device("cuda", 0)
ctx = V.Context(V.make_params(nfeatures=cap, max_width=W, max_height=H, max_batch=F, max_keypoints=cap))
k, a, d, n = ctx.detect_and_compute_batch(np.ascontiguousarray(frames[:, :, :W]), cap)
pts = np.stack([k["x"], k["y"]], -1).astype(np.float32)[:F - 1]
d_f = torch.from_numpy(frames).to(dev)
d_p, d_n = torch.from_numpy(np.ascontiguousarray(pts)).to(dev), torch.from_numpy(n[:F - 1].astype(np.int32)).to(dev)
d_o = torch.zeros(F - 1, cap, 2, dtype=torch.float32, device=dev)
d_s = torch.zeros(F - 1, cap, dtype=torch.uint8, device=dev)
d_e = torch.zeros(F - 1, cap, dtype=torch.float32, device=dev)
torch.cuda.synchronize()


def run():
    ctx.lk_track_batch_ptr(d_f.data_ptr(), F, W, H, PITCH, H * PITCH, d_p.data_ptr(), d_n.data_ptr(), cap, d_o.data_ptr(), d_s.data_ptr(),
                           d_e.data_ptr())
    ctx.synchronize()


run()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ts = []
for _ in range(3):
    torch.cuda.synchronize()
    ev[0].record()
    run()
    ev[1].record()
    torch.cuda.synchronize()
    ts.append(ev[0].elapsed_time(ev[1]))
print("pairs", F - 1, "points", int(n[:F - 1].sum()), "tracked", int(d_s.sum().item()), "ms", round(min(ts), 3),
      "points/s", round(int(n[:F - 1].sum()) / (min(ts) * 1e-3)))
ctx.close()
