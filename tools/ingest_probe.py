"""Wall-clock probe of the file ingest (host decode vs device decode) for a list of wave sizes.
usage: python tools/ingest_probe.py [frames] [real|synth] [chunk ...]"""
import importlib
import os
import shutil
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
V = importlib.import_module("visual-odometry-gpu_b200")

F = int(sys.argv[1]) if len(sys.argv) > 1 else 256
kind = sys.argv[2] if len(sys.argv) > 2 else "synth"
chunks = [int(c) for c in sys.argv[3:]] or [128]
d = tempfile.mkdtemp(prefix="orb_probe_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
try:
    if kind == "real":
        files = [os.path.join(ROOT, "tests", "golden", "kitti_000000.png"), os.path.join(ROOT, "tests", "golden", "kitti_000001.png")]
    else:
        pool = V.synth_frames(32)
        files = []
        for i in range(len(pool)):
            files.append(os.path.join(d, "%06d.png" % i))
            V.synth.write_png_gray8(files[-1], pool[i])
    paths = [files[i % len(files)] for i in range(F)]
    import torch
    KP = V.KP
    outp = (torch.zeros(F, 2000, 2, dtype=torch.int32).pin_memory().numpy().view(KP).reshape(F, 2000),
            torch.zeros(F, 2000, dtype=torch.float32).pin_memory().numpy(),
            torch.zeros(F, 2000, 32, dtype=torch.uint8).pin_memory().numpy(), torch.zeros(F, dtype=torch.int32).pin_memory().numpy())
    for chunk in chunks:
        ctx = V.Context(V.make_params(nfeatures=2000, max_width=1241, max_height=376, max_batch=F, chunk_frames=chunk, max_keypoints=2000))
        res = {}
        for dev in (False, True):
            ctx.detect_and_compute_files(paths, cap=2000, decode_on_device=dev, out=outp)
            t0 = time.perf_counter()
            out = ctx.detect_and_compute_files(paths, cap=2000, decode_on_device=dev, out=outp)
            res[dev] = F / (time.perf_counter() - t0)
        print("frames %d kind %s chunk %d: host decode %.0f fps, device decode %.0f fps, kp %d" % (F, kind, chunk, res[False], res[True], int(out[3].sum())), flush=True)
        ctx.close()
finally:
    shutil.rmtree(d, ignore_errors=True)
