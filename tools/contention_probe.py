"""Do the ORB kernels slow down while the copy engine streams unrelated data into HBM?  (explains the host-staged time)"""
import importlib, os, sys, time, threading
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
V = importlib.import_module("visual-odometry-gpu_b200")
F, W, H, cap = 1000, 1241, 376, 2000
PITCH = (W + 1 + 15) // 16 * 16
dev = torch.device("cuda", 0)
pool = V.synth_frames(64, W, H, pitch=PITCH)
d_frames = torch.from_numpy(pool)[torch.arange(F) % 64].contiguous().to(dev)
db = (torch.zeros(F, cap, 2, dtype=torch.int32, device=dev), torch.zeros(F, cap, dtype=torch.float32, device=dev),
      torch.zeros(F, cap, 32, dtype=torch.uint8, device=dev), torch.zeros(F, dtype=torch.int32, device=dev))
junk_h = torch.empty(256 * 1000 * 1000, dtype=torch.uint8).pin_memory()
junk_d = torch.empty_like(junk_h, device=dev)
junk_h2 = torch.empty(256 * 1000 * 1000, dtype=torch.uint8).pin_memory()
copy_stream = torch.cuda.Stream(device=dev)
copy_stream2 = torch.cuda.Stream(device=dev)
for chunk in (128, 0):
    ctx = V.Context(V.make_params(nfeatures=2000, max_width=W, max_height=H, max_batch=F, chunk_frames=chunk, max_keypoints=cap))
    stream = torch.cuda.Stream(device=dev); ctx.set_stream(stream.cuda_stream)
    def step():
        ctx.detect_and_compute_batch_ptr(d_frames.data_ptr(), 1, F, W, H, PITCH, H * PITCH, cap, db[0].data_ptr(), db[1].data_ptr(), db[2].data_ptr(), db[3].data_ptr(), 1)
    def timed(mode):
        for _ in range(2): step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if mode in ("h2d", "both"):
            with torch.cuda.stream(copy_stream):
                for _ in range(3): junk_d.copy_(junk_h, non_blocking=True)        # ~14 ms of H2D at 55 GB/s
        if mode in ("d2h", "both"):
            with torch.cuda.stream(copy_stream2):
                for _ in range(3): junk_h2.copy_(junk_d, non_blocking=True)
        e0.record(stream)
        step()
        e1.record(stream)
        torch.cuda.synchronize()
        return e0.elapsed_time(e1)
    print("chunk", chunk, {m: round(min(timed(m) for _ in range(3)), 2) for m in ("alone", "h2d", "d2h", "both")}, flush=True)
    ctx.close()
