#!/usr/bin/env python3
"""Small end-to-end exercise of every kernel for compute-sanitizer (memcheck / racecheck):
   compute-sanitizer --tool memcheck python tools/sanitize_run.py"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
V = importlib.import_module("visual-odometry-gpu_b200")

img = V.synth_frames(1, 640, 200)[0]
ctx = V.Context(V.make_params(nfeatures=600, nlevels=5, max_width=640, max_height=200, max_batch=4, chunk_frames=2,
                              keep_side_arrays=1))
k, a, d, npl = ctx.detect_and_compute(img)
print("single:", len(k), list(npl))
frames = V.synth_frames(3, 640, 200)
kk, aa, dd, n = ctx.detect_and_compute_batch(frames)
print("batch:", list(n))
kp = ctx.fast_detect(img, 500)
ang = ctx.orientations(img, kp)
# keypoints in the right / bottom bands: BRIEF boxes that leave the image (decision D7), corner included
ys, xs = np.mgrid[170:197, 610:637]
band = np.zeros(ys.size, V.KP)
band["x"], band["y"] = xs.ravel(), ys.ravel()
rng = np.random.default_rng(1)
des = ctx.brief(img, np.concatenate([kp, band]), np.concatenate([ang, rng.uniform(-3.14, 3.14, len(band)).astype(np.float32)]))
r = ctx.harris(img, kp)
print("stages:", len(kp), des.shape, float(r.max()))
noise = rng.integers(0, 256, (200, 640), dtype=np.uint8)
k2, _, _, npl2 = ctx.detect_and_compute(noise)          # dense corners: exercises the dense fallback of k_fast
print("noise:", len(k2), list(npl2))
ctx.close()
print("SANITIZE_RUN_OK")
