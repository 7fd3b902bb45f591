"""Pyramidal Lucas-Kanade: cv::calcOpticalFlowPyrLK(img_1, img_2, points1, points2, status, err,
Size(21,21), 3, TermCriteria(COUNT+EPS, 30, 0.01), 0, 0.001) of reference src/feature_tracking.cpp:174-180.

OpenCV is a dependency of the reference that is not part of its tree, so the checker (oracle/orb_oracle.cpp, orc_lk_track)
restates OpenCV 4.x lkpyramid.cpp and is pinned here against cv2 4.13 itself: same status for every point, positions within
0.01 px (cv2's SIMD sums the window in another order; p99 of the difference is 1e-4 px, isolated diverging tracks reach 0.03).  The CUDA tracker
sums in the checker's order and has to match it bit for bit.
"""
import importlib
import os

import numpy as np
import pytest

from conftest import GOLDEN

orb = importlib.import_module("visual-odometry-gpu_b200.orb")
CRIT = dict(win=21, max_level=3, max_iter=30, eps=0.01, min_eig=0.001)     # the reference's call


def kitti_pair():
    cv2 = pytest.importorskip("cv2")
    a = cv2.imread(os.path.join(GOLDEN, "kitti_000000.png"), cv2.IMREAD_GRAYSCALE)
    b = cv2.imread(os.path.join(GOLDEN, "kitti_000001.png"), cv2.IMREAD_GRAYSCALE)
    return a, b


def fast_points(img, n):
    import cv2
    kp = cv2.FastFeatureDetector_create(20, True).detect(img)      # as reference src/feature_tracking.cpp:30,59
    return np.array([k.pt for k in kp], np.float32)[:n]


def border_points(w, h):
    return np.array([[0, 0], [w - 1, h - 1], [5.5, 3.25], [w - 2.3, 10.1], [w / 2 + 0.2, h - 1.1], [2, h / 2], [-30, 5], [w + 40, 7],
                     [10.5, 10.5], [w - 11, h - 11], [w - 0.5, 3]], np.float32)


def cv2_track(a, b, pts, win=21, max_level=3, max_iter=30, eps=0.01, min_eig=0.001):
    import cv2
    n1, st, er = cv2.calcOpticalFlowPyrLK(a, b, pts.reshape(-1, 1, 2).copy(), None, winSize=(win, win), maxLevel=max_level,
                                          criteria=(cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, max_iter, eps), flags=0,
                                          minEigThreshold=min_eig)
    return n1.reshape(-1, 2), st.ravel(), er.ravel()


def test_oracle_pyrdown_is_cv2_pyrdown(O):
    cv2 = pytest.importorskip("cv2")
    a, _ = kitti_pair()
    rng = np.random.default_rng(0)
    for img in (a, a[:101, :77], rng.integers(0, 256, (37, 53), dtype=np.uint8), rng.integers(0, 256, (2, 3), dtype=np.uint8)):
        img = np.ascontiguousarray(img)
        assert np.array_equal(O.lk_pyr_down(img), cv2.pyrDown(img))


@pytest.mark.parametrize("cfg", [CRIT, dict(win=15, max_level=2, max_iter=10, eps=0.03, min_eig=1e-4),
                                 dict(win=9, max_level=0, max_iter=30, eps=0.01, min_eig=0.001),
                                 dict(win=31, max_level=5, max_iter=20, eps=0.01, min_eig=0.001)])
def test_oracle_pinned_against_cv2(O, cfg):
    a, b = kitti_pair()
    pts = np.concatenate([fast_points(a, 1500), border_points(a.shape[1], a.shape[0])])
    n1, st1, er1 = cv2_track(a, b, pts, **cfg)
    n2, st2, er2 = O.lk_track(a, b, pts, **cfg)
    assert np.array_equal(st1, st2)
    ok = st1 == 1
    assert ok.mean() > 0.8
    # tolerance: 0.01 px (float summation order inside cv2).  A diverging track (the point ran hundreds of pixels away,
    # err ~ 80) amplifies the last-bit differences: allow 0.5 % of the points up to 0.1 px.
    d = np.abs(n1 - n2).max(1)[ok]
    assert (d < 0.01).mean() > 0.995 and d.max() < 0.1
    assert np.abs(er1 - er2)[ok].max() < 0.1


def test_count_only_criteria_mean_epsilon_001():
    """What include/orb_vo_frontend.hpp substitutes when TermCriteria has no EPS flag: OpenCV then uses epsilon 0.01, whatever
    the epsilon field says (the adapter used 0.001)."""
    import cv2
    a, b = kitti_pair()
    pts = fast_points(a, 800).reshape(-1, 1, 2)

    def run(crit):
        return cv2.calcOpticalFlowPyrLK(a, b, pts.copy(), None, winSize=(21, 21), maxLevel=3, criteria=crit, flags=0, minEigThreshold=0.001)[0]
    count_only = run((cv2.TERM_CRITERIA_COUNT, 30, 0.5))
    assert np.array_equal(count_only, run((cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01)))
    assert not np.array_equal(count_only, run((cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.001)))
    src = open(os.path.join(os.path.dirname(GOLDEN), "..", "include", "orb_vo_frontend.hpp")).read()
    assert "criteria.epsilon : 0.01;" in src


def test_oracle_recovers_a_known_shift(O):
    a, _ = kitti_pair()
    b = np.roll(a, (3, -5), (0, 1))                   # content moves by (+3 rows, -5 columns)
    pts = fast_points(a, 100000)
    pts = pts[(pts[:, 0] > 60) & (pts[:, 0] < a.shape[1] - 60) & (pts[:, 1] > 60) & (pts[:, 1] < a.shape[0] - 60)][::7][:400]
    n2, st2, _ = O.lk_track(a, b, pts, **CRIT)
    d = (n2 - pts)[st2 == 1]
    assert (st2 == 1).mean() > 0.9
    assert np.abs(np.median(d, 0) - np.array([-5, 3])).max() < 0.05


# ---------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("cfg", [CRIT, dict(win=15, max_level=2, max_iter=10, eps=0.03, min_eig=1e-4),
                                 dict(win=9, max_level=0, max_iter=30, eps=0.01, min_eig=0.001),
                                 dict(win=33, max_level=6, max_iter=100, eps=0.0, min_eig=0.0),
                                 dict(win=3, max_level=3, max_iter=5, eps=0.01, min_eig=0.001)])
def test_tracker_bit_exact_against_oracle(O, cfg):
    a, b = kitti_pair()
    pts = np.concatenate([fast_points(a, 3000), border_points(a.shape[1], a.shape[0])])
    ctx = orb.Context(orb.make_params(nfeatures=500, max_width=1241, max_height=376, max_batch=1))
    try:
        n1, st1, er1 = ctx.lk_track(a, b, pts, **cfg)
        n2, st2, er2 = O.lk_track(a, b, pts, **cfg)
        assert np.array_equal(st1, st2)
        assert np.array_equal(n1.view(np.uint32), n2.view(np.uint32))        # bit-exact positions
        assert np.array_equal(er1.view(np.uint32), er2.view(np.uint32))
        top = ctx.lib.orb_lk_levels(a.shape[1], a.shape[0], cfg["win"], cfg["max_level"])
        lvl = a
        for l in range(1, top + 1):
            lvl = O.lk_pyr_down(lvl)
            assert np.array_equal(ctx.lk_get_level(0, l, a.shape[1], a.shape[0]), lvl)
    finally:
        ctx.close()


@pytest.mark.gpu
def test_tracker_shapes_and_noise(O):
    rng = np.random.default_rng(4)
    ctx = orb.Context(orb.make_params(nfeatures=500, max_width=640, max_height=480, max_batch=1))
    try:
        for (h, w) in ((97, 131), (480, 640), (45, 45), (64, 33)):
            a = rng.integers(0, 256, (h, w), dtype=np.uint8)
            a = np.ascontiguousarray((a.astype(np.float32) * 0.3 + np.roll(a, 1, 0) * 0.3 + np.roll(a, 1, 1) * 0.4).astype(np.uint8))
            b = np.roll(a, (1, 2), (0, 1))
            pts = np.concatenate([rng.uniform(-5, max(w, h) + 5, (400, 2)).astype(np.float32), border_points(w, h)])
            n1, st1, er1 = ctx.lk_track(a, b, pts, **CRIT)
            n2, st2, er2 = O.lk_track(a, b, pts, **CRIT)
            assert np.array_equal(st1, st2), (h, w)
            assert np.array_equal(n1.view(np.uint32), n2.view(np.uint32)), (h, w)
            assert np.array_equal(er1.view(np.uint32), er2.view(np.uint32)), (h, w)
        out = ctx.lk_track(a, b, np.zeros((0, 2), np.float32))
        assert len(out[0]) == 0
        with pytest.raises(orb.OrbError):
            ctx.lk_track(a, b, pts, win=35)
    finally:
        ctx.close()


@pytest.mark.gpu
def test_vo_front_end_chain_is_consistent():
    """The pieces either side of the detector, chained the way the reference's two VO loops chain them: PNG files ->
    (device decode) -> ORB -> Hamming 2-NN + ratio test (feature_matching.cpp) against FAST/ORB points followed by LK
    (feature_tracking.cpp).  The two estimates of where a keypoint of frame 0 went in frame 1 must agree."""
    paths = [os.path.join(GOLDEN, "kitti_000000.png"), os.path.join(GOLDEN, "kitti_000001.png")]
    ctx = orb.Context(orb.make_params(nfeatures=2000, max_width=1241, max_height=376, max_batch=2))
    try:
        kps, ang, des, n = ctx.detect_and_compute_files(paths, decode_on_device=True)
        a, b = ctx.get_ingested_frame(0, 1241, 376), ctx.get_ingested_frame(1, 1241, 376)
        k0, k1 = kps[0][:n[0]], kps[1][:n[1]]
        m = ctx.match_knn2(des[0][:n[0]], des[1][:n[1]])
        keep = ctx.ratio_test(m, 0.8)
        assert keep.sum() > 300
        p0 = np.stack([k0["x"], k0["y"]], 1).astype(np.float32)
        p1m = np.stack([k1["x"][m["idx1"]], k1["y"][m["idx1"]]], 1).astype(np.float32)
        p1t, st, err = ctx.lk_track(a, b, p0, **CRIT)
        both = keep & (st == 1) & (err < 15)
        assert both.sum() > 200
        d = np.abs(p1m[both] - p1t[both]).max(1)
        # keypoints are integer pixels on pyramid levels (coordinates scaled by up to 1.2^7): a few pixels of slack
        assert np.median(d) < 1.5 and (d < 4).mean() > 0.8, (np.median(d), (d < 4).mean())
    finally:
        ctx.close()


@pytest.mark.gpu
def test_batch_tracker_is_the_single_pair_tracker(O):
    """orb_lk_track_batch: frame t into frame t + 1 over a sequence, every pyramid built once, one launch for all points.  Host
    form against the checker pair by pair (ragged point counts, one empty pair), then the device-resident form (frames with a
    row pitch, points and results in device memory) against the host form, bit for bit."""
    import torch
    a, b = kitti_pair()
    rng = np.random.default_rng(11)
    frames = np.stack([a, b, np.roll(b, (1, -2), (0, 1)), a, rng.integers(0, 256, a.shape, dtype=np.uint8)])
    F, h, w = frames.shape
    cap = 700
    pts = np.zeros((F - 1, cap, 2), np.float32)
    n = np.array([700, 650, 0, 11], np.int32)
    for p in range(F - 1):
        base = fast_points(frames[p], cap)
        q = np.concatenate([border_points(w, h), base])[:cap]
        pts[p, :len(q)] = q
        n[p] = min(n[p], len(q))
    ctx = orb.Context(orb.make_params(nfeatures=500, max_width=1241, max_height=376, max_batch=F))
    try:
        for cfg in (CRIT, dict(win=9, max_level=0, max_iter=30, eps=0.01, min_eig=0.001)):
            n1, st1, er1 = ctx.lk_track_batch(frames, pts, n, **cfg)
            for p in range(F - 1):
                m = n[p]
                n2, st2, er2 = O.lk_track(frames[p], frames[p + 1], pts[p, :m], **cfg)
                assert np.array_equal(st1[p, :m], st2), p
                assert np.array_equal(n1[p, :m].view(np.uint32), n2.view(np.uint32)), p
                assert np.array_equal(er1[p, :m].view(np.uint32), er2.view(np.uint32)), p
                assert not st1[p, m:].any()
        # all pairs full (n_pts = None) == per-pair calls
        n1, st1, er1 = ctx.lk_track_batch(frames[:3], pts[:2], None, **CRIT)
        for p in range(2):
            n2, st2, er2 = ctx.lk_track(frames[p], frames[p + 1], pts[p], **CRIT)
            assert np.array_equal(st1[p], st2) and np.array_equal(n1[p].view(np.uint32), n2.view(np.uint32))
        # device-resident: pitched frames, device points / counts / results
        pitch = (w + 1 + 15) // 16 * 16
        pf = np.zeros((F, h, pitch), np.uint8)
        pf[:, :, :w] = frames
        dev = torch.device("cuda", 0)
        d_f = torch.from_numpy(pf).to(dev)
        d_p, d_n = torch.from_numpy(pts).to(dev), torch.from_numpy(n).to(dev)
        d_o = torch.zeros(F - 1, cap, 2, dtype=torch.float32, device=dev)
        d_s = torch.zeros(F - 1, cap, dtype=torch.uint8, device=dev)
        d_e = torch.zeros(F - 1, cap, dtype=torch.float32, device=dev)
        torch.cuda.synchronize()                         # the context works on its own stream
        ctx.lk_track_batch_ptr(d_f.data_ptr(), F, w, h, pitch, h * pitch, d_p.data_ptr(), d_n.data_ptr(), cap, d_o.data_ptr(),
                               d_s.data_ptr(), d_e.data_ptr(), **CRIT)
        ctx.synchronize()
        n1, st1, er1 = ctx.lk_track_batch(frames, pts, n, **CRIT)
        assert np.array_equal(d_s.cpu().numpy(), st1)
        assert np.array_equal(d_o.cpu().numpy().view(np.uint32), n1.view(np.uint32))
        assert np.array_equal(d_e.cpu().numpy().view(np.uint32), er1.view(np.uint32))
        assert ctx.lk_track_batch(frames[:1], np.zeros((0, cap, 2), np.float32))[0].shape == (0, cap, 2)
    finally:
        ctx.close()
