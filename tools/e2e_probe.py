"""Where the host-staged (e2e) time goes: the four combinations of host / device frames and results.
usage: python tools/e2e_probe.py [chunk ...]"""
import importlib, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
V = importlib.import_module("visual-odometry-gpu_b200")
F, W, H, cap = 1000, 1241, 376, 2000
PITCH = (W + 1 + 15) // 16 * 16
pool = V.synth_frames(64, W, H, pitch=PITCH)
h_frames = torch.from_numpy(pool)[torch.arange(F) % 64].contiguous().pin_memory()
dev = torch.device("cuda", 0)
d_frames = h_frames.to(dev)
def bufs(pin):
    mk = (lambda *a, **k: torch.zeros(*a, **k).pin_memory()) if pin else (lambda *a, **k: torch.zeros(*a, device=dev, **k))
    return mk(F, cap, 2, dtype=torch.int32), mk(F, cap, dtype=torch.float32), mk(F, cap, 32, dtype=torch.uint8), mk(F, dtype=torch.int32)
hb, db = bufs(True), bufs(False)
for chunk in [int(c) for c in sys.argv[1:]] or [0]:
    ctx = V.Context(V.make_params(nfeatures=2000, max_width=W, max_height=H, max_batch=F, chunk_frames=chunk, max_keypoints=cap))
    stream = torch.cuda.Stream(device=dev); torch.cuda.set_stream(stream); ctx.set_stream(stream.cuda_stream)
    res = {}
    for name, fr, fdev, ob, odev in (("dev->dev", d_frames, 1, db, 1), ("host->dev", h_frames, 0, db, 1), ("dev->host", d_frames, 1, hb, 0), ("host->host", h_frames, 0, hb, 0)):
        def step():
            ctx.detect_and_compute_batch_ptr(fr.data_ptr(), fdev, F, W, H, PITCH, H * PITCH, cap, ob[0].data_ptr(), ob[1].data_ptr(), ob[2].data_ptr(), ob[3].data_ptr(), odev)
        for _ in range(3): step()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(6): step()
        torch.cuda.synchronize(); res[name] = (time.perf_counter() - t0) / 6 * 1e3
    print("chunk", chunk, {k: round(v, 2) for k, v in res.items()}, flush=True)
    ctx.close()
