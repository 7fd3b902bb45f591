"""Frame ingest: PNG -> 8-bit gray, the replacement of cv::imread(path, IMREAD_GRAYSCALE)
(reference src/feature_matching.cpp:55,59; src/feature_tracking.cpp:56,196).

CPU tests: the host decoder against (a) PNGs assembled here with Python's zlib (every filter type, every deflate
block type, odd IDAT splits) whose pixels are known by construction and (b) cv2 4.13's imread/imdecode.
GPU tests: the file pipeline (decode threads -> pinned staging -> waves) gives the results of the batch call on
the decoded frames.
"""
import os
import struct
import zlib

import numpy as np
import pytest

import importlib

from conftest import GOLDEN as GOLDEN_DIR

orb = importlib.import_module("visual-odometry-gpu_b200.orb")


def synth_frames(n, w, h):
    return importlib.import_module("visual-odometry-gpu_b200.synth").synth_frames(n, w, h)


def chunk(tag, data):
    return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data))


def paeth(a, b, c):
    p = a + b - c
    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
    return a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)


def filter_rows(img, bpp, filters):
    """img: (H, rowbytes) uint8; returns the filtered scanlines (PNG spec 9.2) with the filter byte in front."""
    H, rb = img.shape
    out = bytearray()
    prev = np.zeros(rb, np.int64)
    for y in range(H):
        cur = img[y].astype(np.int64)
        ft = filters[y % len(filters)]
        left = np.concatenate([np.zeros(bpp, np.int64), cur[:-bpp]]) if rb > bpp else np.zeros(rb, np.int64)
        upleft = np.concatenate([np.zeros(bpp, np.int64), prev[:-bpp]]) if rb > bpp else np.zeros(rb, np.int64)
        if ft == 0:
            f = cur
        elif ft == 1:
            f = cur - left
        elif ft == 2:
            f = cur - prev
        elif ft == 3:
            f = cur - ((left + prev) >> 1)
        else:
            f = cur - np.array([paeth(int(a), int(b), int(c)) for a, b, c in zip(left, prev, upleft)], np.int64)
        out.append(ft)
        out += bytes((f & 0xff).astype(np.uint8))
        prev = cur
    return bytes(out)


def make_png(img, color_type=0, depth=8, filters=(0,), level=6, strategy=zlib.Z_DEFAULT_STRATEGY, idat_split=None, interlace=0):
    """img: (H, W[, C]) uint8/uint16 array in PNG channel order."""
    a = np.asarray(img)
    H, W = a.shape[:2]
    ch = 1 if a.ndim == 2 else a.shape[2]
    raw = a.astype(">u2").tobytes() if depth == 16 else a.astype(np.uint8).tobytes()
    bpp = ch * depth // 8
    rows = np.frombuffer(raw, np.uint8).reshape(H, W * bpp)
    co = zlib.compressobj(level, zlib.DEFLATED, 15, 8, strategy)
    z = co.compress(filter_rows(rows, bpp, filters)) + co.flush()
    ihdr = struct.pack(">IIBBBBB", W, H, depth, color_type, 0, 0, interlace)
    out = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", ihdr) + chunk(b"tEXt", b"Comment\x00synthetic")
    if idat_split is None:
        out += chunk(b"IDAT", z)
    else:
        pos = 0
        k = 0
        while pos < len(z):
            n = idat_split[k % len(idat_split)]
            out += chunk(b"IDAT", z[pos:pos + n])
            pos += n
            k += 1
    return out + chunk(b"IEND", b"")


def test_exports_and_info():
    data = open(os.path.join(GOLDEN_DIR, "kitti_000000.png"), "rb").read()
    assert orb.png_info(data) == (1241, 376, 8, 1)
    assert orb.png_info(data[:40]) == (1241, 376, 8, 1)      # the file head is enough
    with pytest.raises(orb.OrbError):
        orb.png_info(b"\x89PNG\r\n\x1a\n" + b"\0" * 40)


@pytest.mark.parametrize("filters", [(0,), (1,), (2,), (3,), (4,), (4, 3, 2, 1, 0), (2, 4)])
def test_every_filter_type_gray8(filters):
    rng = np.random.default_rng(len(filters) * 7 + filters[0])
    img = rng.integers(0, 256, (23, 57), dtype=np.uint8)
    img[5:12, 10:40] = 200                                   # flat area: long matches
    out = orb.imdecode_gray8(make_png(img, filters=filters))
    assert np.array_equal(out, img)


@pytest.mark.parametrize("level,strategy", [(0, zlib.Z_DEFAULT_STRATEGY), (1, zlib.Z_DEFAULT_STRATEGY), (9, zlib.Z_DEFAULT_STRATEGY),
                                            (6, zlib.Z_FIXED), (6, zlib.Z_HUFFMAN_ONLY), (6, zlib.Z_RLE), (6, zlib.Z_FILTERED)])
def test_every_deflate_block_type(level, strategy):
    rng = np.random.default_rng(level + 10 * strategy)
    # smooth ramp + noise + constant stripes: literals, short and maximal matches, far distances (rows are 1500 bytes)
    x = np.arange(1500)
    img = ((x[None, :] // 7 + np.arange(64)[:, None] * 3) % 256).astype(np.uint8)
    img[::5] = rng.integers(0, 256, (13, 1500), dtype=np.uint8)
    img[20:30] = 17
    out = orb.imdecode_gray8(make_png(img, filters=(0, 1, 4), level=level, strategy=strategy, idat_split=(1, 8192, 3, 100)))
    assert np.array_equal(out, img)


def test_large_random_and_empty_like_shapes():
    rng = np.random.default_rng(3)
    for shape in ((1, 1), (1, 300), (300, 1), (2, 2), (376, 1241)):
        img = rng.integers(0, 256, shape, dtype=np.uint8)
        assert np.array_equal(orb.imdecode_gray8(make_png(img, filters=(4, 1, 3))), img)
    assert np.array_equal(orb.imdecode_gray8(make_png(np.zeros((400, 2000), np.uint8))), np.zeros((400, 2000), np.uint8))


def test_other_layouts_reduce_like_libpng():
    rng = np.random.default_rng(5)
    g16 = rng.integers(0, 65536, (9, 31), dtype=np.uint16)
    assert np.array_equal(orb.imdecode_gray8(make_png(g16, 0, 16, filters=(4,))), (g16 >> 8).astype(np.uint8))
    ga = rng.integers(0, 256, (9, 31, 2), dtype=np.uint8)
    assert np.array_equal(orb.imdecode_gray8(make_png(ga, 4, 8, filters=(3,))), ga[..., 0])
    rgb = rng.integers(0, 256, (9, 31, 3), dtype=np.uint8)
    rgb[0, :5] = 77                                           # r == g == b passes through
    r, g, b = [rgb[..., i].astype(np.int64) for i in range(3)]
    want = ((9797 * r + 19234 * g + 3737 * b) >> 15).astype(np.uint8)
    assert np.array_equal(orb.imdecode_gray8(make_png(rgb, 2, 8, filters=(4, 1))), want)
    rgba = np.concatenate([rgb, rng.integers(0, 256, (9, 31, 1), dtype=np.uint8)], -1)
    assert np.array_equal(orb.imdecode_gray8(make_png(rgba, 6, 8, filters=(2,))), want)


def test_rejections():
    img = np.arange(64, dtype=np.uint8).reshape(8, 8)
    good = make_png(img, filters=(1,))
    assert np.array_equal(orb.imdecode_gray8(good), img)
    def with_ihdr(depth=8, color=0, interlace=0):
        ihdr = struct.pack(">IIBBBBB", 8, 8, depth, color, 0, 0, interlace)
        return good[:8] + chunk(b"IHDR", ihdr) + good[8 + 25:]

    assert with_ihdr() == good
    for bad, what in [(with_ihdr(interlace=1), "interlaced"), (with_ihdr(color=3), "palette"),
                      (with_ihdr(depth=4), "bit depth"), (good[:len(good) // 2], "truncated"), (good[:-12], "IEND")]:
        with pytest.raises(orb.OrbError) as e:
            orb.imdecode_gray8(bad)
        assert e.value.code == -7, what
    # any single bit flip in IHDR / IDAT is caught by the chunk CRC (or the zlib checksum)
    rng = np.random.default_rng(0)
    idat_at = good.index(b"IDAT")
    for _ in range(40):
        pos = int(rng.integers(idat_at + 4, len(good) - 16))
        bad = bytearray(good)
        bad[pos] ^= 1 << int(rng.integers(0, 8))
        with pytest.raises(orb.OrbError):
            orb.imdecode_gray8(bytes(bad))
    # a stream that is valid PNG framing around corrupt deflate data must fail cleanly, never crash
    for _ in range(200):
        z = bytearray(zlib.compress(bytes(rng.integers(0, 8, 600, dtype=np.uint8))))
        z[int(rng.integers(2, len(z)))] ^= 1 << int(rng.integers(0, 8))
        png = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", 24, 24, 8, 0, 0, 0, 0)) + chunk(b"IDAT", bytes(z)) + chunk(b"IEND", b"")
        try:
            orb.imdecode_gray8(png)
        except orb.OrbError:
            pass


def test_against_cv2():
    cv2 = pytest.importorskip("cv2")
    for name in ("kitti_000000.png", "kitti_000001.png"):
        p = os.path.join(GOLDEN_DIR, name)
        assert np.array_equal(orb.imread_gray8(p), cv2.imread(p, cv2.IMREAD_GRAYSCALE))
    base = cv2.imread(os.path.join(GOLDEN_DIR, "kitti_000000.png"), cv2.IMREAD_GRAYSCALE)
    rng = np.random.default_rng(1)
    col = rng.integers(0, 256, (97, 131, 3), dtype=np.uint8)
    col16 = rng.integers(0, 65536, (97, 131, 3), dtype=np.uint16)
    cases = [(base, [cv2.IMWRITE_PNG_COMPRESSION, c, cv2.IMWRITE_PNG_STRATEGY, s])
             for c in (0, 1, 9) for s in (cv2.IMWRITE_PNG_STRATEGY_DEFAULT, cv2.IMWRITE_PNG_STRATEGY_RLE, cv2.IMWRITE_PNG_STRATEGY_FIXED)]
    cases += [(np.stack([np.roll(base, 5, 1), base, np.roll(base, 3, 0)], -1), []), (col, []), (col16, []),
              (np.concatenate([col, col[..., :1]], -1), []), (base.astype(np.uint16) * 257 + 13, [])]
    for img, params in cases:
        ok, buf = cv2.imencode(".png", img, params)
        assert ok
        assert np.array_equal(orb.imdecode_gray8(buf.tobytes()), cv2.imdecode(buf, cv2.IMREAD_GRAYSCALE))


# ---------------------------------------------------------------------------------------------
def _write_frames(tmp_path, frames, **kw):
    paths = []
    for i, f in enumerate(frames):
        p = os.path.join(str(tmp_path), "%06d.png" % i)
        with open(p, "wb") as fh:
            fh.write(make_png(f, filters=(1, 4, 2), **kw))
        paths.append(p)
    return paths


def _raw_deflate(data, level, strategy):
    co = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
    return co.compress(data) + co.flush()


@pytest.mark.gpu
def test_device_inflate_against_zlib():
    """k_inflate on raw deflate streams of every block type; the expected bytes are the inputs of Python's zlib."""
    rng = np.random.default_rng(11)
    datas, streams = [], []
    for trial in range(160):
        kind = trial % 8
        n = int(rng.integers(0, 120000)) if trial % 16 else int(rng.integers(0, 40))
        if kind == 0:
            d = rng.integers(0, 256, n, dtype=np.uint8)
        elif kind == 1:
            d = rng.integers(0, 4, n, dtype=np.uint8)
        elif kind == 2:
            d = (np.arange(n) % 251).astype(np.uint8)
        elif kind == 3:
            d = np.zeros(n, np.uint8)
        elif kind == 4:
            d = np.repeat(rng.integers(0, 256, n // 50 + 1, dtype=np.uint8), 50)[:n]
        elif kind == 5:
            d = rng.normal(128, 6, n).clip(0, 255).astype(np.uint8)
        elif kind == 6:   # far matches: a block repeated at distances up to 32 KiB
            blk = rng.integers(0, 256, int(rng.integers(100, 33000)), dtype=np.uint8)
            d = np.tile(blk, n // len(blk) + 2)[:n]
        else:             # skewed alphabet: long Huffman codes (second-level tables)
            d = np.minimum(rng.geometric(0.35, n), 255).astype(np.uint8)
        level = int(rng.integers(0, 10))
        strat = int(rng.choice([zlib.Z_DEFAULT_STRATEGY, zlib.Z_FILTERED, zlib.Z_HUFFMAN_ONLY, zlib.Z_RLE, zlib.Z_FIXED]))
        datas.append(d.tobytes())
        streams.append(_raw_deflate(datas[-1], level, strat))
    p = orb.make_params(nfeatures=500, max_width=320, max_height=200, max_batch=4)
    ctx = orb.Context(p)
    try:
        outs, st = ctx.debug_inflate(streams, [len(d) for d in datas])
        assert not st.any(), st
        for k, (o, d) in enumerate(zip(outs, datas)):
            assert o == d, k
        # wrong expected size, truncated and bit-flipped streams end with a status (or, for a flip that still decodes
        # to the right size, with some output): they never hang
        big = [k for k in range(len(datas)) if len(datas[k]) > 2000][:24]
        bad_streams, sizes = [], []
        for k in big:
            s = bytearray(streams[k])
            s[int(rng.integers(0, len(s)))] ^= 1 << int(rng.integers(0, 8))
            bad_streams += [bytes(s), streams[k][:len(streams[k]) // 2], streams[k]]
            sizes += [len(datas[k]), len(datas[k]), len(datas[k]) - 1]
        outs, st = ctx.debug_inflate(bad_streams, sizes)
        assert all(st[3 * j + 1] != 0 and st[3 * j + 2] != 0 for j in range(len(big))), st
    finally:
        ctx.close()


@pytest.mark.gpu
def test_device_decode_all_filter_types(tmp_path):
    rng = np.random.default_rng(2)
    frames = synth_frames(5, 320, 200)
    frames[3] = rng.integers(0, 256, (200, 320), dtype=np.uint8)
    frames[4, 50:120] = 9
    paths = []
    for i, f in enumerate(frames):
        p = os.path.join(str(tmp_path), "%d.png" % i)
        open(p, "wb").write(make_png(f, filters=[(4,), (3,), (2, 1), (0, 4, 3, 2, 1), (4, 4, 1, 3)][i], level=[6, 1, 9, 0, 6][i],
                                     idat_split=[None, (8192,), (1, 77), (5000,), (3,)][i]))
        paths.append(p)
    p = orb.make_params(nfeatures=500, max_width=320, max_height=200, max_batch=8)
    ctx = orb.Context(p)
    try:
        want = ctx.detect_and_compute_batch(frames)
        got = ctx.detect_and_compute_files(paths, decode_on_device=True)
        for f in range(5):
            assert np.array_equal(ctx.get_ingested_frame(f, 320, 200), frames[f]), f
        assert np.array_equal(got[3], want[3]) and np.array_equal(got[2], want[2]) and np.array_equal(got[0], want[0])
    finally:
        ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("threads,device", [(1, False), (0, False), (0, True)])
def test_files_pipeline_equals_batch(tmp_path, threads, device):
    frames = synth_frames(37, 1241, 376)
    paths = []
    for i, f in enumerate(frames):
        paths.append(os.path.join(str(tmp_path), "%06d.png" % i))
        importlib.import_module("visual-odometry-gpu_b200.synth").write_png_gray8(paths[-1], f, level=[6, 1, 9][i % 3])
    paths[5] = os.path.join(GOLDEN_DIR, "kitti_000000.png")          # a real KITTI file (Huffman-only stream)
    import cv2
    frames[5] = cv2.imread(paths[5], cv2.IMREAD_GRAYSCALE)
    p = orb.make_params(nfeatures=2000, max_width=1241, max_height=376, max_batch=64, chunk_frames=8)
    ctx = orb.Context(p)
    try:
        want = ctx.detect_and_compute_batch(frames)
        got = ctx.detect_and_compute_files(paths, threads=threads, decode_on_device=device)
        for f in (0, 17, 36):
            assert np.array_equal(ctx.get_ingested_frame(f, 1241, 376), frames[f])
        assert np.array_equal(got[3], want[3])
        for f in range(len(frames)):
            n = int(want[3][f])
            assert np.array_equal(got[0][f, :n], want[0][f, :n])
            assert np.array_equal(got[1][f, :n].view(np.uint32), want[1][f, :n].view(np.uint32))
            assert np.array_equal(got[2][f, :n], want[2][f, :n])
    finally:
        ctx.close()


@pytest.mark.gpu
def test_files_pipeline_errors(tmp_path):
    frames = synth_frames(6, 320, 200)
    paths = _write_frames(tmp_path, frames)
    p = orb.make_params(nfeatures=500, max_width=320, max_height=200, max_batch=16)
    ctx = orb.Context(p)
    try:
        with pytest.raises(orb.OrbError) as e:
            ctx.detect_and_compute_files(paths[:3] + [os.path.join(str(tmp_path), "missing.png")] + paths[3:])
        assert e.value.code == -6 and "missing.png" in str(e.value)
        bad = os.path.join(str(tmp_path), "bad.png")
        data = bytearray(open(paths[2], "rb").read())
        data[len(data) // 2] ^= 0x10
        open(bad, "wb").write(bytes(data))
        with pytest.raises(orb.OrbError) as e:
            ctx.detect_and_compute_files(paths[:2] + [bad])
        assert e.value.code == -7 and "bad.png" in str(e.value)
        other = os.path.join(str(tmp_path), "other.png")
        open(other, "wb").write(make_png(np.zeros((100, 100), np.uint8)))
        with pytest.raises(orb.OrbError):
            ctx.detect_and_compute_files(paths[:2] + [other])
        # device decode: a stream that is corrupt inside valid chunk CRCs is reported by the kernel's status
        z = bytearray(zlib.compress(filter_rows(frames[0], 1, (1,)), 6))
        z[len(z) // 3] ^= 0x04
        crafted = os.path.join(str(tmp_path), "crafted.png")
        open(crafted, "wb").write(b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", 320, 200, 8, 0, 0, 0, 0)) +
                                  chunk(b"IDAT", bytes(z)) + chunk(b"IEND", b""))
        with pytest.raises(orb.OrbError) as e:
            ctx.detect_and_compute_files(paths[:2] + [crafted] + paths[2:], decode_on_device=True)
        assert e.value.code == -7 and "crafted.png" in str(e.value)
        # ... and a damaged file is caught by the chunk CRCs, which the device checks (k_png_crc): a flipped data byte and a
        # flipped byte of the stored CRC itself both fail, like they do in libpng
        with pytest.raises(orb.OrbError) as e:
            ctx.detect_and_compute_files(paths[:2] + [bad], decode_on_device=True)
        assert e.value.code == -7 and "bad.png" in str(e.value) and "CRC" in str(e.value)
        data = bytearray(open(paths[3], "rb").read())
        at = data.rindex(b"IEND") - 8          # last byte of the last IDAT chunk's CRC field
        data[at] ^= 0x01
        crcflip = os.path.join(str(tmp_path), "crcflip.png")
        open(crcflip, "wb").write(bytes(data))
        for dev in (False, True):
            with pytest.raises(orb.OrbError) as e:
                ctx.detect_and_compute_files([crcflip] + paths[:2], decode_on_device=dev)
            assert e.value.code == -7 and "CRC" in str(e.value)
        rgb = os.path.join(str(tmp_path), "rgb.png")
        open(rgb, "wb").write(make_png(np.zeros((200, 320, 3), np.uint8), color_type=2))
        with pytest.raises(orb.OrbError) as e:
            ctx.detect_and_compute_files(paths[:2] + [rgb], decode_on_device=True)
        assert e.value.code == -7
        # the context stays usable
        got = ctx.detect_and_compute_files(paths, decode_on_device=True)
        want = ctx.detect_and_compute_batch(frames)
        assert np.array_equal(got[3], want[3]) and np.array_equal(got[2], want[2])
        got = ctx.detect_and_compute_files(paths)
        want = ctx.detect_and_compute_batch(frames)
        assert np.array_equal(got[3], want[3]) and np.array_equal(got[2], want[2])
    finally:
        ctx.close()
