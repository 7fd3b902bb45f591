"""The C++ facade (include/orb.hpp, include/orb_cpu.hpp: the reference's class names and signatures) compiled with
g++ against liborb_b200.so and driven the way the reference's src/compare.cpp drives its classes."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build(tmp_path, V):
    exe = str(tmp_path / "facade_test")
    libdir = os.path.dirname(V.lib_path())
    cmd = ["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "facade_test.cpp"),
           "-o", exe, "-L" + libdir, "-lorb_b200", "-Wl,-rpath," + libdir]
    env = {k: v for k, v in os.environ.items() if k not in ("CXX", "CC")}
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    assert r.returncode == 0, r.stdout
    return exe


def test_facade_compiles_without_gpu(tmp_path, V):
    """Header-only facade + C ABI link on the CPU box (no compute call)."""
    V.load_library()
    _build(tmp_path, V)


@pytest.mark.gpu
def test_facade_matches_oracle(tmp_path, V, O, kitti0):
    V.load_library()
    exe = _build(tmp_path, V)
    raw = tmp_path / "k0.raw"
    kitti0.tofile(raw)
    out = str(tmp_path / "out")
    r = subprocess.run([exe, str(raw), "1241", "376", out], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0 and "FACADE_OK" in r.stdout, r.stdout
    assert "pattern_sum %d" % int(O.pattern().astype(int).sum()) in r.stdout

    def load(tag):
        k = np.fromfile(out + "." + tag + ".kps", dtype=V.KP)
        a = np.fromfile(out + "." + tag + ".ang", dtype=np.float32)
        d = np.fromfile(out + "." + tag + ".desc", dtype=np.uint8).reshape(-1, 32)
        return k, a, d

    # ORB() == reference constructor defaults 500 / 1.2 / 8 (thr 20, patch 31, Harris top-N)
    k, a, d = load("orb")
    ref = O.detect_and_compute(kitti0, O.params(), cap=500)
    assert np.array_equal(k, ref["kps"]) and np.array_equal(a.view(np.uint32), ref["angles"].view(np.uint32))
    assert np.array_equal(d, ref["desc"])
    # ORBCPU() == the reference's shipped CPU path (single level, thr 50, raster 3000, patch 9)
    k, a, d = load("cpu")
    ref = O.detect_and_compute(kitti0, O.params(nfeatures=3000, nlevels=1, fast_threshold=50, orient_patch=9, select_policy=0), cap=3000)
    assert len(k) == 1178 and np.array_equal(k, ref["kps"]) and np.array_equal(d, ref["desc"])
    # stage classes
    k, a, d = load("stage")
    kr = O.nms(O.fast_scores(kitti0, 20), 3, 700)
    assert np.array_equal(k, kr)
    ar = O.orientations(kitti0, kr, 31)
    assert np.array_equal(a.view(np.uint32), ar.view(np.uint32)) and np.array_equal(d, O.brief(kitti0, kr, ar))
    # the free functions of include/orb_stages.hpp (reference Fast.cuh / Brief.cuh / HarrisScore.cuh)
    k, a, d = load("free")
    assert np.array_equal(k, kr) and np.array_equal(a.view(np.uint32), ar.view(np.uint32)) and np.array_equal(d, O.brief(kitti0, kr, ar))
    hs = np.fromfile(out + ".free.harris", dtype=np.float32)
    assert np.array_equal(hs.view(np.uint32), O.harris(kitti0, kr, 0.04).view(np.uint32))
    h0 = np.fromfile(out + ".free.harris_k0", dtype=np.float32)                     # the reference's `int k` signature: k = 0
    assert np.array_equal(h0.view(np.uint32), O.harris(kitti0, kr, 0.0).view(np.uint32))
    # the reference's stand-alone filters through the facade
    G5 = np.float32([1, 4, 7, 4, 1, 4, 16, 26, 16, 4, 7, 26, 41, 26, 7, 4, 16, 26, 16, 4, 1, 4, 7, 4, 1])

    def filt(tag, shape=(376, 1241)):
        return np.fromfile(out + ".filt." + tag, dtype=np.uint8).reshape(shape)
    assert np.array_equal(filt("g5"), O.conv2d_u8(kitti0, G5, True, 273.0))
    assert np.array_equal(filt("g1d"), O.gaussian_blur_1d(kitti0))
    assert np.array_equal(filt("g7"), O.conv2d_u8(kitti0, O.gaussian_kernel(7), True))
    assert np.array_equal(filt("sx"), O.conv2d_u8(kitti0, np.float32([-1, 0, 1, -2, 0, 2, -1, 0, 1]), True))
    assert np.array_equal(filt("sy"), O.conv2d_u8(kitti0, np.float32([-1, -2, -1, 0, 0, 0, 1, 2, 1]), True))
    assert np.array_equal(filt("c3", (374, 1239)), O.conv2d_u8(kitti0, np.float32([0.111, 0.111, 0.111, 0.111, 0.112, 0.111, 0.111, 0.111, 0.111])))
    # NMS() over the synthetic score map of facade_test.cpp
    ys, xs = np.mgrid[0:376, 0:1241]
    m = (((xs * 7 + ys * 13) % 31) * ((xs ^ ys) & 1)).astype(np.float32)
    pad = np.zeros((378, 1243), np.float32)
    pad[1:-1, 1:-1] = m
    best = np.max([pad[1 + dy:377 + dy, 1 + dx:1242 + dx] for dy in (-1, 0, 1) for dx in (-1, 0, 1)], axis=0)
    keep = (m > 10) & (m >= best)
    keep[0, :] = keep[-1, :] = False
    keep[:, 0] = keep[:, -1] = False
    ky, kx = np.nonzero(keep)
    kn = np.fromfile(out + ".free.nms", dtype=V.KP)
    assert np.array_equal(kn["x"], kx[:5000]) and np.array_equal(kn["y"], ky[:5000])


def _build_feature2d(tmp_path, V):
    exe = str(tmp_path / "feature2d_test")
    libdir = os.path.dirname(V.lib_path())
    cmd = ["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "tests", "cpp", "mock_opencv"), "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "feature2d_test.cpp"), "-o", exe, "-L" + libdir, "-lorb_b200", "-Wl,-rpath," + libdir]
    env = {k: v for k, v in os.environ.items() if k not in ("CXX", "CC")}
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    assert r.returncode == 0, r.stdout
    return exe


def test_feature2d_adapter_compiles_against_mock_opencv(tmp_path, V):
    """the cv::Feature2D adapter (include/orb_feature2d.hpp) against the mock OpenCV headers."""
    V.load_library()
    _build_feature2d(tmp_path, V)


@pytest.mark.gpu
def test_feature2d_adapter_matches_oracle(tmp_path, V, O, kitti0):
    V.load_library()
    exe = _build_feature2d(tmp_path, V)
    raw = tmp_path / "k0.raw"
    kitti0.tofile(raw)
    out = str(tmp_path / "f2d")
    r = subprocess.run([exe, str(raw), "1241", "376", out], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0 and "FEATURE2D_OK" in r.stdout, r.stdout
    kp = np.fromfile(out + ".kp", dtype=np.float32).reshape(-1, 6)
    desc = np.fromfile(out + ".desc", dtype=np.uint8).reshape(-1, 32)
    ref = O.detect_and_compute(kitti0, O.params(nfeatures=3000), cap=3000)          # ORB::create(3000): 1.2 / 8 / thr 20 / patch 31
    assert len(kp) == ref["n"] and np.array_equal(desc, ref["desc"])
    assert np.array_equal(kp[:, 0], ref["kps"]["x"].astype(np.float32)) and np.array_equal(kp[:, 1], ref["kps"]["y"].astype(np.float32))
    assert np.array_equal(kp[:, 5].astype(np.int32), ref["level_id"])
    assert np.array_equal(kp[:, 4].view(np.uint32), ref["response"].view(np.uint32))
    deg = ref["angles"] * np.float32(180.0 / np.pi)
    deg = np.where(deg < 0, deg + np.float32(360), deg)
    assert np.allclose(kp[:, 3], deg, atol=1e-4) and (kp[:, 3] >= 0).all() and (kp[:, 3] < 360).all()
    size = np.float32(31) * np.float32(1.2) ** ref["level_id"].astype(np.float32)
    assert np.allclose(kp[:, 2], size, rtol=1e-6)


def _build_frontend(tmp_path, V):
    exe = str(tmp_path / "frontend_test")
    libdir = os.path.dirname(V.lib_path())
    cmd = ["g++", "-std=c++17", "-O1", "-I" + os.path.join(ROOT, "tests", "cpp", "mock_opencv"), "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "tests", "cpp", "frontend_test.cpp"), "-o", exe, "-L" + libdir, "-lorb_b200", "-Wl,-rpath," + libdir]
    env = {k: v for k, v in os.environ.items() if k not in ("CXX", "CC")}
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    assert r.returncode == 0, r.stdout
    return exe


def test_vo_frontend_compiles_against_mock_opencv(tmp_path, V):
    """orb_b200::imread / orb_b200::calcOpticalFlowPyrLK (include/orb_vo_frontend.hpp)."""
    V.load_library()
    _build_frontend(tmp_path, V)


@pytest.mark.gpu
def test_vo_frontend_matches_oracle(tmp_path, V, O, kitti0, kitti1):
    V.load_library()
    exe = _build_frontend(tmp_path, V)
    g = os.path.join(ROOT, "tests", "golden")
    ys, xs = np.mgrid[20:360:17, 20:1220:23]
    pts = np.stack([xs.ravel(), ys.ravel()], 1).astype(np.float32) + np.float32(0.25)
    pts.tofile(tmp_path / "pts.f32")
    out = str(tmp_path / "out")
    r = subprocess.run([exe, os.path.join(g, "kitti_000000.png"), os.path.join(g, "kitti_000001.png"), str(tmp_path / "pts.f32"), out],
                       stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0 and "FRONTEND_OK 1241 x 376" in r.stdout, r.stdout
    assert np.array_equal(np.fromfile(out + ".img", np.uint8).reshape(376, 1241), kitti0)      # orb_b200::imread == cv2.imread
    n2, st2, er2 = O.lk_track(kitti0, kitti1, pts, win=21, max_level=3, max_iter=30, eps=0.01, min_eig=0.001)
    assert np.array_equal(np.fromfile(out + ".st", np.uint8), st2)
    assert np.array_equal(np.fromfile(out + ".pts", np.uint32), n2.view(np.uint32).ravel())
    assert np.array_equal(np.fromfile(out + ".err", np.uint32), er2.view(np.uint32))
