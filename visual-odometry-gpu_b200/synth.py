"""Synthetic KITTI-like frames: frame i = the fixture frame mirror-tiled to WxH, circularly
rolled by (17*i mod H, 113*i mod W), plus i.i.d. integer noise uniform[-3,3] from PCG64(seed=i), clipped to u8.
Corner density stays close to the real frame (about 3 % of pixels pass FAST at threshold 20)."""
import os

import numpy as np

_FIXTURE = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "kitti_000000.png")
_base_cache = {}


def load_fixture(path=None):
    import cv2
    path = path or _FIXTURE
    img = cv2.imread(path, cv2.IMREAD_GRAYSCALE)
    if img is None:
        raise FileNotFoundError(path)
    return img


def _base(W, H):
    key = (W, H)
    if key not in _base_cache:
        img = load_fixture()
        h, w = img.shape
        ty, tx = -(-H // h), -(-W // w)
        rows = []
        for j in range(ty):
            r = img if j % 2 == 0 else img[::-1]
            rows.append(np.concatenate([r if i % 2 == 0 else r[:, ::-1] for i in range(tx)], axis=1))
        _base_cache[key] = np.ascontiguousarray(np.concatenate(rows, axis=0)[:H, :W])
    return _base_cache[key]


def synth_frame(i, W=1241, H=376):
    base = _base(W, H)
    f = np.roll(base, ((17 * i) % H, (113 * i) % W), axis=(0, 1)).astype(np.int16)
    noise = np.random.Generator(np.random.PCG64(i)).integers(-3, 4, size=(H, W), dtype=np.int16)
    return np.clip(f + noise, 0, 255).astype(np.uint8)


def synth_frames(n, W=1241, H=376, start=0, pitch=None):
    """(n, H, pitch) uint8 array (pitch >= W; columns >= W are zero)."""
    pitch = pitch or W
    out = np.zeros((n, H, pitch), np.uint8)
    for k in range(n):
        out[k, :, :W] = synth_frame(start + k, W, H)
    return out


def write_png_gray8(path, img, level=6, huffman_only=False):
    """Minimal 8-bit gray PNG writer for the ingest bench / tests (Sub filter on every row, like the KITTI files,
    IDAT chunks of 8 KB like libpng writes them; huffman_only=True gives match-free deflate streams, which is what the
    KITTI odometry files contain).  Test-data generator, not part of the product path."""
    import struct
    import zlib
    a = np.ascontiguousarray(img, np.uint8)
    H, W = a.shape
    rows = np.empty((H, W + 1), np.uint8)
    rows[:, 0] = 1
    rows[:, 1] = a[:, 0]
    rows[:, 2:] = a[:, 1:] - a[:, :-1]
    if huffman_only:
        co = zlib.compressobj(level, zlib.DEFLATED, 15, 8, zlib.Z_HUFFMAN_ONLY)
        z = co.compress(rows.tobytes()) + co.flush()
    else:
        z = zlib.compress(rows.tobytes(), level)

    def chunk(tag, data):
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data))
    out = [b"\x89PNG\r\n\x1a\n", chunk(b"IHDR", struct.pack(">IIBBBBB", W, H, 8, 0, 0, 0, 0))]
    out += [chunk(b"IDAT", z[i:i + 8192]) for i in range(0, len(z), 8192)]
    out.append(chunk(b"IEND", b""))
    with open(path, "wb") as f:
        f.write(b"".join(out))
