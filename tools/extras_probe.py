"""Small run of the kernels outside the ORB step (ingest: k_inflate, k_unfilter; tracker: k_lk_pyrdown, k_lk_track) for an
ncu launch list: python tools/extras_probe.py [frames]"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
V = importlib.import_module("visual-odometry-gpu_b200")
F = int(sys.argv[1]) if len(sys.argv) > 1 else 256
g = os.path.join(ROOT, "tests", "golden")
files = [os.path.join(g, "kitti_000000.png"), os.path.join(g, "kitti_000001.png")]
ctx = V.Context(V.make_params(nfeatures=2000, max_width=1241, max_height=376, max_batch=F, max_keypoints=2000))
out = ctx.detect_and_compute_files([files[i % 2] for i in range(F)], cap=2000, decode_on_device=True)
f0, f1 = V.imread_gray8(files[0]), V.imread_gray8(files[1])
kp = out[0][0][:int(out[3][0])]
pts = np.stack([kp["x"], kp["y"]], 1).astype(np.float32)
for _ in range(3):
    nxt, st, err = ctx.lk_track(f0, f1, pts)
print("frames", F, "keypoints", int(out[3].sum()), "tracked", int(st.sum()), "of", len(pts))
ctx.close()
