"""Does running two halves of a batch on two streams (two contexts, two scratch arenas) beat one stream?  usage: two_stream_probe.py [frames] [chunk]"""
import importlib, sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
V = importlib.import_module("visual-odometry-gpu_b200")
F = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 0
W, H, P, cap = 1241, 376, 1248, 2000
dev = torch.device("cuda", 0)
frames = torch.from_numpy(V.synth_frames(F, W, H, pitch=P)).to(dev)
def mk(n):
    return (torch.zeros(n, cap, 2, dtype=torch.int32, device=dev), torch.zeros(n, cap, dtype=torch.float32, device=dev),
            torch.zeros(n, cap, 32, dtype=torch.uint8, device=dev), torch.zeros(n, dtype=torch.int32, device=dev))
def ctx(n):
    return V.Context(V.make_params(nfeatures=2000, nlevels=8, max_width=W, max_height=H, max_batch=n, max_keypoints=cap, chunk_frames=chunk))
def run(parts):
    cs = [ctx(n) for _, n in parts]; outs = [mk(n) for _, n in parts]; ss = [torch.cuda.Stream() for _ in parts]
    for c, s in zip(cs, ss): c.set_stream(s.cuda_stream)
    def step():
        for (o, n), c, out in zip(parts, cs, outs):
            c.detect_and_compute_batch_ptr(frames[o:o + n].data_ptr(), 1, n, W, H, P, H * P, cap, out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(), out[3].data_ptr(), 1)
    for _ in range(3): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in ss: s.wait_event(e0)
    for _ in range(10): step()
    for s in ss: torch.cuda.current_stream().wait_stream(s)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    tot = sum(int(o[3].sum()) for o in outs)
    for c in cs: c.close()
    return ms, tot
for parts in ([(0, F)], [(0, F // 2), (F // 2, F - F // 2)], [(0, F // 4), (F // 4, F // 4), (F // 2, F // 4), (3 * F // 4, F - 3 * (F // 4))]):
    ms, tot = run(parts)
    print("%d stream(s): %.3f ms per %d frames = %.0f frames/s (kp %d)" % (len(parts), ms, F, F / ms * 1e3, tot))
