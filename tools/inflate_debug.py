"""Device inflate vs zlib on the deflate streams of synthetic frames (diagnostic)."""
import importlib, os, sys, zlib
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
V = importlib.import_module("visual-odometry-gpu_b200")
frames = V.synth_frames(32)
datas, streams = [], []
for f in frames:
    rows = np.empty((f.shape[0], f.shape[1] + 1), np.uint8)
    rows[:, 0] = 1; rows[:, 1] = f[:, 0]; rows[:, 2:] = f[:, 1:] - f[:, :-1]
    raw = rows.tobytes()
    for level, mem in ((6, 8), (6, 9), (1, 8)):
        co = zlib.compressobj(level, zlib.DEFLATED, -15, mem)
        streams.append(co.compress(raw) + co.flush()); datas.append(raw)
ctx = V.Context(V.make_params(nfeatures=500, max_width=320, max_height=200, max_batch=4))
outs, st = ctx.debug_inflate(streams, [len(d) for d in datas])
print("status", st.tolist())
for k, (o, d) in enumerate(zip(outs, datas)):
    if o != d:
        a = np.frombuffer(o, np.uint8); b = np.frombuffer(d, np.uint8)
        bad = np.nonzero(a != b)[0]
        print("stream", k, "status", st[k], "first mismatch at", int(bad[0]), "of", len(d), "count", len(bad), "last", int(bad[-1]))
print("done")
