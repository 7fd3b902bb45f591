"""Parity tests proper: the CUDA path, called through the C ABI (ctypes -> liborb_b200.so), against the CPU oracle
and the committed golden fixtures.  Integer work (pyramid pixels, keypoint sets, BRIEF bits) must be bit-exact; the
float work (Harris response, orientation) is ALSO required bit-exact here (tolerance 0 ulp), because the device
code restates the reference's IEEE operation sequences including glibc's atan2f / sinf / cosf."""
import numpy as np
import pytest

from conftest import noise_image

pytestmark = pytest.mark.gpu


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def ctx_kitti(V):
    """8 levels, N=2000, orb.hpp defaults (thr 20, patch 31), Harris top-N, side arrays on."""
    c = V.Context(V.make_params(nfeatures=2000, nlevels=8, threshold=20, patch_size=31, max_width=1241, max_height=376,
                                max_batch=8, keep_side_arrays=1, max_keypoints=2000))
    yield c
    c.close()


def test_library_is_the_cuda_one(V):
    import torch
    assert torch.cuda.is_available()
    assert "liborb_b200.so" in V.lib_path()
    maps = open("/proc/self/maps").read()
    V.load_library()
    assert "liborb_b200.so" in open("/proc/self/maps").read() or "liborb_b200.so" in maps


# ------------------------------------------------------------------------------------------------- libm twins
def test_device_libm_twins_bit_exact(V, O, ctx_kitti):
    rng = np.random.default_rng(0)
    n = 200000
    # moments are integers: |m| <= 1.9e6 for patch 31
    y = rng.integers(-1900000, 1900001, n).astype(np.float32)
    x = rng.integers(-1900000, 1900001, n).astype(np.float32)
    y[:2000] = rng.integers(-300, 301, 2000)
    x[:2000] = rng.integers(-300, 301, 2000)
    y[2000:2100] = 0
    x[2100:2200] = 0
    x[2200:2300] = 1
    got = ctx_kitti.eval_math(0, y, x)
    L = O.lib()
    ref = np.array([L.orc_atan2f(float(a), float(b)) for a, b in zip(y[:30000], x[:30000])], np.float32)
    assert np.array_equal(bits(got[:30000]), bits(ref))
    ang = np.concatenate([got[:30000], rng.uniform(-np.pi, np.pi, 30000).astype(np.float32),
                          rng.uniform(-1e-3, 1e-3, 2000).astype(np.float32), np.float32([0, np.pi, -np.pi, np.pi / 4, -np.pi / 4])])
    c, s = ctx_kitti.eval_math(1, ang), ctx_kitti.eval_math(2, ang)
    cr = np.array([L.orc_cosf(float(a)) for a in ang], np.float32)
    sr = np.array([L.orc_sinf(float(a)) for a in ang], np.float32)
    assert np.array_equal(bits(c), bits(cr)) and np.array_equal(bits(s), bits(sr))
    v = np.concatenate([rng.uniform(-30, 30, 50000).astype(np.float32), (np.arange(-80, 81) / 2).astype(np.float32),
                        np.float32([0.49999997, -0.49999997, 0.5, -0.5, 1.5, 2.5, -2.5])])
    r = ctx_kitti.eval_math(3, v)
    rr = np.array([L.orc_lround_f(float(a)) for a in v], np.float32)
    assert np.array_equal(r, rr)


# ------------------------------------------------------------------------------------------------- pyramid
@pytest.mark.parametrize("shape", [(376, 1241), (370, 1226), (200, 333)])
@pytest.mark.parametrize("blur", [1, 0])
def test_pyramid_pixels_bit_exact(V, O, kitti0, shape, blur):
    h, w = shape
    img = kitti0[:h, :w] if shape != (200, 333) else noise_image(h, w, 21)
    c = V.Context(V.make_params(nfeatures=500, nlevels=8, blur_levels=blur, max_width=w, max_height=h))
    c.detect_and_compute(img)
    p = O.params(nlevels=8, blur_levels=blur)
    for l in range(8):
        got = c.get_level(0, l, w, h)
        ref = O.build_level(img, p, l)
        assert got.shape == ref.shape
        assert np.array_equal(got, ref), "level %d differs in %d px" % (l, int((got != ref).sum()))
    c.close()


def test_pyramid_every_width_residue(V, O):
    """Every width mod 16 (row padding 1..16 bytes, so the last source words of a row sit at every distance from the
    end of the pitch) and tile edges that fall anywhere: both resize paths of k_pyramid (column pairs from word loads
    while a level shrinks by <= 3, byte gathers below) against the oracle, with and without the blur."""
    for w in range(129, 161):
        h = 67 + (w % 5)
        img = noise_image(h, w, 1000 + w)
        for blur in (1, 0):
            c = V.Context(V.make_params(nfeatures=200, nlevels=8, blur_levels=blur, max_width=w, max_height=h))
            c.detect_and_compute(img)
            p = O.params(nlevels=8, blur_levels=blur)
            for l in range(8):
                got, ref = c.get_level(0, l, w, h), O.build_level(img, p, l)
                assert got.shape == ref.shape
                assert np.array_equal(got, ref), "w %d blur %d level %d differs in %d px" % (w, blur, l, int((got != ref).sum()))
            c.close()


def test_pyramid_matches_cv2_golden_hashes(V, golden, kitti0, kitti1, ctx_kitti):
    import hashlib
    for tag, img in (("k0", kitti0), ("k1", kitti1)):
        ctx_kitti.detect_and_compute(img)
        for l in range(1, 8):
            lv = ctx_kitti.get_level(0, l, 1241, 376)
            assert hashlib.sha256(lv.tobytes()).hexdigest() == str(golden["cv2_%s_pyr_sha" % tag][l - 1])


# ------------------------------------------------------------------------------------------------- FAST + NMS
@pytest.mark.parametrize("thr,n,nms", [(20, 9, 3), (50, 9, 3), (20, 12, 3), (35, 9, 1), (0, 9, 3), (250, 9, 3)])
def test_fast_nms_keypoint_sets(V, O, kitti0, thr, n, nms):
    c = V.Context(V.make_params(nfeatures=3000, nlevels=1, threshold=thr, n=n, nms_window=nms,
                                select_policy=V.SELECT_RASTER_FIRST_N, max_width=1241, max_height=376))
    for img in (kitti0, noise_image(123, 517, 3, "blocks"), noise_image(90, 140, 4)):
        ref_all = O.nms(O.fast_scores(img, thr, n), nms)
        for cap in (3000, 100, 0):
            got = c.fast_detect(img, cap)
            assert np.array_equal(got, ref_all[:cap]), (thr, n, nms, cap, len(got), len(ref_all))
    c.close()


def test_nms_over_a_callers_score_map(V, O, kitti0):
    """NMS() of the reference's stage seam (include/NMS.cuh:5, src/cuda/NMS.cu:95-127): score > threshold, no strictly greater
    score in the window (ties keep both), centres at least window/2 inside the map; survivors in raster order."""
    c = V.Context(V.make_params(nfeatures=3000, nlevels=1, select_policy=V.SELECT_RASTER_FIRST_N, max_width=1241, max_height=376))
    sc = O.fast_scores(kitti0, 20).astype(np.float32)
    assert np.array_equal(c.nms_scores(sc, 3, 8192), O.nms(sc, 3))                 # FAST scores: the FAST + NMS stage itself
    assert np.array_equal(c.nms_scores(sc, 3, 100), O.nms(sc, 3)[:100])
    rng = np.random.default_rng(11)
    for (h, w), win, thr in (((60, 90), 3, 0.5), ((47, 131), 5, 0.25), ((33, 40), 1, 0.9), ((20, 300), 7, 0.0)):
        m = rng.random((h, w), dtype=np.float32)
        m[rng.random((h, w)) < 0.3] = np.float32(0.75)                                 # plateaus: exact ties
        r = win // 2
        pad = np.zeros((h + 2 * r, w + 2 * r), np.float32)
        pad[r:h + r, r:w + r] = m
        best = np.max([pad[r + dy:r + dy + h, r + dx:r + dx + w] for dy in range(-r, r + 1) for dx in range(-r, r + 1)], axis=0)
        keep = (m > thr) & (m >= best)
        keep[:r, :] = keep[h - r:, :] = False
        keep[:, :r] = keep[:, w - r:] = False
        ys, xs = np.nonzero(keep)
        got = c.nms_scores(np.ascontiguousarray(np.pad(m, ((0, 0), (0, 5)))[:, :w]), win, 8192, thr)     # padded rows (pitch > w)
        assert len(got) == len(xs) > 0 and np.array_equal(got["x"], xs) and np.array_equal(got["y"], ys)
    c.close()


def test_reference_filter_wrappers(V, O, kitti0):
    """conv2d / GaussianBlur / GaussianBlur1D / GaussianBlurCUDA / SobelCUDA of the reference's stage seam (include/Convolution.cuh:5,
    GaussianBlur.cuh:3-4, GaussianBlur.hpp:6, Sobel.hpp:6): float accumulation with one FMA per tap, convertTo(CV_8U); byte-exact
    against the oracle's restatement, which is itself tied to cv2 where OpenCV has the same operator."""
    import cv2
    c = V.Context(V.make_params(nlevels=1, max_width=1241, max_height=376))
    rng = np.random.default_rng(3)
    G5 = np.float32([1, 4, 7, 4, 1, 4, 16, 26, 16, 4, 7, 26, 41, 26, 7, 4, 16, 26, 16, 4, 1, 4, 7, 4, 1])
    for img in (kitti0, noise_image(37, 53, 2), noise_image(5, 7, 3)):
        assert np.array_equal(c.conv2d_u8(img, G5, True, 273.0), O.conv2d_u8(img, G5, True, 273.0))              # GaussianBlur
        assert np.array_equal(c.gaussian_blur_1d(img), O.gaussian_blur_1d(img))                                      # GaussianBlur1D
        for ks in (3, 5, 7):
            assert np.array_equal(c.conv2d_u8(img, O.gaussian_kernel(ks), True), O.conv2d_u8(img, O.gaussian_kernel(ks), True))   # GaussianBlurCUDA
        for k in (np.float32([-1, 0, 1, -2, 0, 2, -1, 0, 1]), np.float32([-1, -2, -1, 0, 0, 0, 1, 2, 1])):           # SobelCUDA
            got = c.conv2d_u8(img, k, True)
            assert np.array_equal(got, O.conv2d_u8(img, k, True))
            ref = cv2.Sobel(img, cv2.CV_32F, int(k[2] == 1), int(k[2] != 1), ksize=3, borderType=cv2.BORDER_REFLECT_101)
            assert np.array_equal(got, np.clip(ref, 0, 255).astype(np.uint8))
        kr = rng.normal(0, 0.3, 25).astype(np.float32)                                                                # conv2d, valid mode
        if img.shape[0] >= 5:
            assert np.array_equal(c.conv2d_u8(img, kr), O.conv2d_u8(img, kr))
    c.close()


def test_reference_single_level_mode_golden(V, golden, kitti0, kitti1):
    """ORBCPU as shipped (D3) -- compared with the outputs of the reference's own compiled code."""
    orb = V.ORBCPU(max_width=1241, max_height=376)
    for tag, img in (("k0", kitti0), ("k1", kitti1)):
        k, a, d = orb.detectAndCompute(img)
        assert np.array_equal(k, golden["ref_%s_t50_kps" % tag])
        assert np.array_equal(bits(a), bits(golden["ref_%s_t50_ang" % tag]))
        assert np.array_equal(d, golden["ref_%s_t50_desc" % tag])
    orb20 = V.ORB(nfeatures=3000, nlevels=1, threshold=20, patch_size=31, select_policy=V.SELECT_RASTER_FIRST_N,
                  max_width=1241, max_height=376, max_keypoints=3000)
    for tag, img in (("k0", kitti0), ("k1", kitti1)):
        k, a, d = orb20.detectAndCompute(img)
        assert np.array_equal(k, golden["ref_%s_t20_kps" % tag])
        assert np.array_equal(bits(a), bits(golden["ref_%s_t20_ang" % tag]))
        assert np.array_equal(d, golden["ref_%s_t20_desc" % tag])


# ------------------------------------------------------------------------------------------------- Harris
def test_harris_stage_bit_exact(V, O, kitti0, ctx_kitti):
    assert np.array_equal(bits(ctx_kitti.harris_weights()), bits(O.harris_weights()))
    for img in (kitti0, noise_image(101, 203, 8), noise_image(64, 64, 9, "blocks")):
        k = O.nms(O.fast_scores(img, 20))
        # plus border-hugging points (x,y = 3 / dim-4) that exercise the reflect-101 Sobel taps
        h, w = img.shape
        extra = np.array([(3, 3), (w - 4, 3), (3, h - 4), (w - 4, h - 4), (w // 2, 3), (3, h // 2)], dtype=V.KP)
        k = np.concatenate([k, extra])
        got, ref = ctx_kitti.harris(img, k), O.harris(img, k, 0.04)
        assert np.array_equal(bits(got), bits(ref)), int((bits(got) != bits(ref)).sum())


def test_candidates_and_responses_per_level(V, O, kitti0, ctx_kitti):
    ctx_kitti.detect_and_compute(kitti0)
    p = O.params(nlevels=8)
    for l in range(8):
        lv = O.build_level(kitti0, p, l)
        ref = O.nms(O.fast_scores(lv, 20))
        xy, rsp, n = ctx_kitti.get_candidates(0, l)
        assert n == len(ref) == len(xy)
        order = np.argsort(xy["y"].astype(np.int64) * 65536 + xy["x"])
        assert np.array_equal(xy[order], ref)                                        # keypoint SETS equal (D4)
        assert np.array_equal(bits(rsp[order]), bits(O.harris(lv, ref, 0.04)))       # Harris bit-exact (D5)


# ------------------------------------------------------------------------------------------------- orientation / BRIEF
@pytest.mark.parametrize("patch", [31, 9, 15])
def test_orientation_stage_bit_exact(V, O, kitti0, patch):
    c = V.Context(V.make_params(nlevels=1, patch_size=patch, max_width=1241, max_height=376))
    for img in (kitti0, noise_image(77, 130, 12)):
        h, w = img.shape
        k = O.nms(O.fast_scores(img, 20))
        got, ref = c.orientations(img, k), O.orientations(img, k, patch)
        assert np.array_equal(bits(got), bits(ref)), int((bits(got) != bits(ref)).sum())
        assert (ref == 0).any() or patch < 9                                         # some border keypoints -> 0.0f
    c.close()


def test_brief_stage_bit_exact_including_border_rule(V, O, kitti0, ctx_kitti):
    rng = np.random.default_rng(5)
    for img in (kitti0, noise_image(80, 97, 13), noise_image(41, 60, 14)):
        h, w = img.shape
        k = O.nms(O.fast_scores(img, 20))
        # every pixel of the right / bottom bands where the reference's sum5x5 leaves the integral image (D7)
        ys, xs = np.mgrid[max(3, h - 24):h - 3, max(3, w - 24):w - 3]
        band = np.zeros(ys.size, V.KP)
        band["x"], band["y"] = xs.ravel(), ys.ravel()
        right = np.zeros(40, V.KP)
        right["x"], right["y"] = rng.integers(max(3, w - 20), w - 3, 40), rng.integers(3, h - 3, 40)
        bottom = np.zeros(40, V.KP)
        bottom["x"], bottom["y"] = rng.integers(3, w - 3, 40), rng.integers(max(3, h - 20), h - 3, 40)
        topleft = np.zeros(20, V.KP)
        topleft["x"], topleft["y"] = rng.integers(3, 25, 20), rng.integers(3, 25, 20)
        k = np.concatenate([k, band, right, bottom, topleft])
        ang = np.concatenate([O.orientations(img, k[:len(k) // 2], 31),
                              rng.uniform(-np.pi, np.pi, len(k) - len(k) // 2).astype(np.float32)])
        fl = O.brief_flags(w, h, k, ang)
        assert (fl & 2).any() and (fl & 1).any()                 # the undefined-read and skip paths are exercised
        got, ref = ctx_kitti.brief(img, k, ang), O.brief(img, k, ang)
        bad = np.flatnonzero((got != ref).any(axis=1))
        assert len(bad) == 0, (len(bad), k[bad[:5]], fl[bad[:5]])


# ------------------------------------------------------------------------------------------------- whole path
def _compare_full(ctx, O, img, p, cap):
    k, a, d, npl = ctx.detect_and_compute(img, cap)
    r = O.detect_and_compute(img, p, cap=cap)
    assert len(k) == r["n"] and np.array_equal(npl, r["n_per_level"])
    assert np.array_equal(k, r["kps"])
    assert np.array_equal(bits(a), bits(r["angles"]))
    ident = (d == r["desc"]).all(axis=1)
    assert ident.all(), "descriptor identity %.4f" % ident.mean()
    xy, lid, rsp = ctx.get_side_arrays(0, len(k))
    assert np.array_equal(xy, r["level_xy"]) and np.array_equal(lid, r["level_id"])
    assert np.array_equal(bits(rsp), bits(r["response"]))
    return len(k)


def test_full_path_fixture_frames_vs_oracle_and_golden(V, O, golden, kitti0, kitti1, ctx_kitti):
    p = O.params(nfeatures=2000, nlevels=8, fast_threshold=20, orient_patch=31, select_policy=1)
    for tag, img in (("k0", kitti0), ("k1", kitti1)):
        n = _compare_full(ctx_kitti, O, img, p, 2000)
        assert n == 1996
        k, a, d, npl = ctx_kitti.detect_and_compute(img, 2000)
        assert np.array_equal(k, golden["ml_%s_t20_kps" % tag]) and np.array_equal(d, golden["ml_%s_t20_desc" % tag])
        assert np.array_equal(bits(a), bits(golden["ml_%s_t20_angles" % tag]))


def test_full_path_orb_cpu_parameter_set(V, O, golden, kitti0):
    """thr 50 / patch 9 (include/orb_cpu.hpp defaults): upper levels have fewer candidates than quota."""
    c = V.Context(V.make_params(nfeatures=2000, nlevels=8, threshold=50, patch_size=9, max_width=1241, max_height=376,
                                keep_side_arrays=1, max_keypoints=2000))
    p = O.params(nfeatures=2000, nlevels=8, fast_threshold=50, orient_patch=9, select_policy=1)
    _compare_full(c, O, kitti0, p, 2000)
    k, a, d, npl = c.detect_and_compute(kitti0, 2000)
    assert np.array_equal(npl, golden["ml_k0_t50_n_per_level"]) and np.array_equal(d, golden["ml_k0_t50_desc"])
    c.close()


@pytest.mark.parametrize("kind,shape,N,L", [("uniform", (120, 200), 500, 4), ("blocks", (150, 260), 300, 5),
                                            ("uniform", (61, 67), 200, 8), ("uniform", (376, 1241), 2000, 8)])
def test_full_path_synthetic_and_ties(V, O, kind, shape, N, L):
    """Noise (dense corners, ~10x real density) and flat blocks (exact score / response ties)."""
    h, w = shape
    img = noise_image(h, w, 31, kind)
    c = V.Context(V.make_params(nfeatures=N, nlevels=L, max_width=w, max_height=h, keep_side_arrays=1))
    p = O.params(nfeatures=N, nlevels=L)
    _compare_full(c, O, img, p, c.max_kp)
    c.close()


def test_raster_policy_multilevel_and_caps(V, O, kitti0):
    c = V.Context(V.make_params(nfeatures=300, nlevels=3, select_policy=V.SELECT_RASTER_FIRST_N, max_width=1241,
                                max_height=376, keep_side_arrays=1))
    p = O.params(nfeatures=300, nlevels=3, select_policy=0)
    _compare_full(c, O, kitti0, p, 900)
    k, a, d, npl = c.detect_and_compute(kitti0, 450)         # output cap truncates in level order
    r = O.detect_and_compute(kitti0, p, cap=450)
    assert len(k) == 450 and np.array_equal(k, r["kps"]) and np.array_equal(d, r["desc"])
    c.close()


def test_degenerate_frames(V, O):
    c = V.Context(V.make_params(nfeatures=100, nlevels=3, max_width=64, max_height=64))
    k, a, d, npl = c.detect_and_compute(np.full((40, 50), 128, np.uint8))
    assert len(k) == 0 and not npl.any()
    tiny = noise_image(9, 11, 5)
    k, a, d, npl = c.detect_and_compute(tiny)
    r = O.detect_and_compute(tiny, O.params(nfeatures=100, nlevels=3), cap=100)
    assert len(k) == r["n"] and np.array_equal(k, r["kps"]) and np.array_equal(d, r["desc"])
    with pytest.raises(V.OrbError) as e:
        c.detect_and_compute(np.zeros((65, 64), np.uint8))
    assert e.value.code == -3
    with pytest.raises(ValueError):
        c.detect_and_compute(np.zeros((10, 10), np.float32))
    c.close()


def test_batch_chunking_and_strided_rows(V, O, ctx_kitti):
    """A batch gives per-frame results identical to single calls, whatever the chunk size and row pitch."""
    frames = V.synth_frames(5, 1241, 376)
    p = O.params(nfeatures=2000, nlevels=8)
    n_ref, k_ref, a_ref, d_ref = O.detect_and_compute_batch(frames, p, 2000, 4, keep=True)
    for chunk in (0, 2, 8):
        c = V.Context(V.make_params(nfeatures=2000, nlevels=8, max_width=1241, max_height=376, max_batch=8,
                                    chunk_frames=chunk, max_keypoints=2000))
        k, a, d, n = c.detect_and_compute_batch(frames, 2000)
        assert np.array_equal(n, n_ref)
        for f in range(5):
            m = n[f]
            assert np.array_equal(k[f, :m], k_ref[f, :m]) and np.array_equal(d[f, :m], d_ref[f, :m])
            assert np.array_equal(bits(a[f, :m]), bits(a_ref[f, :m]))
        c.close()
    padded = V.synth_frames(3, 1241, 376, pitch=1280)[:, :, :1241]        # non-contiguous rows (pitch 1280)
    k, a, d, n = ctx_kitti.detect_and_compute_batch(padded, 2000)
    assert np.array_equal(n[:3], n_ref[:3]) and np.array_equal(d[2, :n[2]], d_ref[2, :n[2]])


def test_device_resident_batch_through_torch_pointers(V, O):
    """The bench path: frames and outputs live in torch CUDA tensors, work runs on torch's current stream."""
    import torch
    F, W, H, cap, pitch = 6, 1241, 376, 2000, 1248
    frames = V.synth_frames(F, W, H, pitch=pitch)
    c = V.Context(V.make_params(nfeatures=2000, nlevels=8, max_width=W, max_height=H, max_batch=F, chunk_frames=4))
    d_frames = torch.from_numpy(frames).cuda()
    d_k = torch.zeros(F, cap, 2, dtype=torch.int32, device="cuda")
    d_a = torch.zeros(F, cap, dtype=torch.float32, device="cuda")
    d_d = torch.zeros(F, cap, 32, dtype=torch.uint8, device="cuda")
    d_n = torch.zeros(F, dtype=torch.int32, device="cuda")
    c.set_stream(torch.cuda.current_stream().cuda_stream)
    c.detect_and_compute_batch_ptr(d_frames.data_ptr(), 1, F, W, H, pitch, H * pitch, cap, d_k.data_ptr(),
                                   d_a.data_ptr(), d_d.data_ptr(), d_n.data_ptr(), 1)
    c.synchronize()
    p = O.params(nfeatures=2000, nlevels=8)
    n_ref, k_ref, a_ref, d_ref = O.detect_and_compute_batch(np.ascontiguousarray(frames[:, :, :W]), p, cap, 4, keep=True)
    n = d_n.cpu().numpy()
    assert np.array_equal(n, n_ref)
    k = d_k.cpu().numpy()
    dd = d_d.cpu().numpy()
    for f in range(F):
        assert np.array_equal(k[f, :n[f], 0], k_ref[f, :n[f]]["x"]) and np.array_equal(k[f, :n[f], 1], k_ref[f, :n[f]]["y"])
        assert np.array_equal(dd[f, :n[f]], d_ref[f, :n[f]])
    assert c.launch_count() == 2 * 6                                       # 6 kernels x 2 chunks
    c.close()


@pytest.mark.parametrize("W", [175, 176 - 3, 161])
def test_device_frames_with_minimal_row_padding(V, O, W):
    """Device-resident frames whose pitch leaves 1, 3 or 15 spare bytes per row are read in place (no staging copy):
    the word loads of k_pyramid / k_harris must stay inside each row."""
    import torch
    F, H, cap, pitch = 3, 90, 600, 176
    rng = np.random.default_rng(W)
    frames = rng.integers(0, 256, (F, H, pitch), dtype=np.uint8)
    c = V.Context(V.make_params(nfeatures=500, nlevels=6, max_width=W, max_height=H, max_batch=F))
    d_frames = torch.from_numpy(frames).cuda()
    d_k = torch.zeros(F, cap, 2, dtype=torch.int32, device="cuda")
    d_a = torch.zeros(F, cap, dtype=torch.float32, device="cuda")
    d_d = torch.zeros(F, cap, 32, dtype=torch.uint8, device="cuda")
    d_n = torch.zeros(F, dtype=torch.int32, device="cuda")
    c.set_stream(torch.cuda.current_stream().cuda_stream)
    c.detect_and_compute_batch_ptr(d_frames.data_ptr(), 1, F, W, H, pitch, H * pitch, cap, d_k.data_ptr(),
                                   d_a.data_ptr(), d_d.data_ptr(), d_n.data_ptr(), 1)
    c.synchronize()
    p = O.params(nfeatures=500, nlevels=6)
    n_ref, k_ref, a_ref, d_ref = O.detect_and_compute_batch(np.ascontiguousarray(frames[:, :, :W]), p, cap, 2, keep=True)
    n, k, a, dd = d_n.cpu().numpy(), d_k.cpu().numpy(), d_a.cpu().numpy(), d_d.cpu().numpy()
    assert np.array_equal(n, n_ref) and n.min() > 100
    for f in range(F):
        assert np.array_equal(k[f, :n[f], 0], k_ref[f, :n[f]]["x"]) and np.array_equal(k[f, :n[f], 1], k_ref[f, :n[f]]["y"])
        assert np.array_equal(bits(a[f, :n[f]]), bits(a_ref[f, :n[f]]))
        assert np.array_equal(dd[f, :n[f]], d_ref[f, :n[f]])
    c.close()


def test_reference_class_surface(V, O, kitti0):
    """OrientedFAST / RotatedBRIEF / ORB with the reference's constructor defaults (include/orb.hpp:10-49)."""
    fast = V.OrientedFAST(max_width=1241, max_height=376)
    k = fast.detect(kitti0, 2 * 434)                                       # ORB::detectAndCompute asks 2*quota
    ref = O.nms(O.fast_scores(kitti0, 20), 3, 2 * 434)
    assert np.array_equal(k, ref)
    a = fast.compute_orientations(kitti0, k)
    assert np.array_equal(bits(a), bits(O.orientations(kitti0, k, 31)))
    d = V.RotatedBRIEF(max_width=1241, max_height=376).compute(kitti0, k, a)
    assert np.array_equal(d, O.brief(kitti0, k, a))
    orb = V.ORB(max_width=1241, max_height=376)                            # 500 / 1.2 / 8
    kk, aa, dd = orb.detectAndCompute(kitti0)
    r = O.detect_and_compute(kitti0, O.params(), cap=500)
    assert np.array_equal(kk, r["kps"]) and np.array_equal(dd, r["desc"])


# ------------------------------------------------------------------------------------------------- BASELINE configs 4 / 5
@pytest.mark.parametrize("W,H,L,N", [(1920, 1080, 8, 5000), (3840, 2160, 12, 10000)])
def test_full_path_large_frames(V, O, W, H, L, N):
    """BASELINE.json configs 4 and 5 (1080p / 8 levels / 5000 kp; 4K / 12 levels / 10000 kp): one synthetic frame each,
    every output compared with the oracle (the oracle needs a few seconds per frame at these sizes)."""
    img = V.synth_frames(1, W, H)[0]
    c = V.Context(V.make_params(nfeatures=N, nlevels=L, max_width=W, max_height=H, keep_side_arrays=1))
    p = O.params(nfeatures=N, nlevels=L)
    n = _compare_full(c, O, img, p, c.max_kp)
    assert n == sum(O.level_quota(N, 1.2, L, l) for l in range(L))           # every level fills its quota
    for l in (1, L - 1):
        assert np.array_equal(c.get_level(0, l, W, H), O.build_level(img, p, l))
    c.close()


def test_batch_properties_at_bench_size(V, O):
    """Size-independent properties on a bench-sized batch (200 frames, 2 waves): identical frames give identical
    records; per-frame results do not depend on the position in the batch or on the wave size; every frame of the
    synthetic set fills its budget; one frame is spot-checked against the oracle."""
    F, W, H, cap = 200, 1241, 376, 2000
    pool = V.synth_frames(8, W, H)
    frames = pool[np.arange(F) % 8]
    c = V.Context(V.make_params(nfeatures=2000, nlevels=8, max_width=W, max_height=H, max_batch=F, max_keypoints=cap))
    k, a, d, n = c.detect_and_compute_batch(frames, cap)
    assert (n == 1996).all()
    for f in range(8, F):
        assert np.array_equal(k[f], k[f % 8]) and np.array_equal(d[f], d[f % 8]) and np.array_equal(a[f], a[f % 8])
    c2 = V.Context(V.make_params(nfeatures=2000, nlevels=8, max_width=W, max_height=H, max_batch=F, max_keypoints=cap,
                                 chunk_frames=7))
    k2, a2, d2, n2 = c2.detect_and_compute_batch(frames[::-1].copy(), cap)
    assert np.array_equal(k2[::-1], k) and np.array_equal(d2[::-1], d)
    r = O.detect_and_compute(pool[5], O.params(nfeatures=2000, nlevels=8), cap=cap)
    assert np.array_equal(k[5, :n[5]], r["kps"]) and np.array_equal(d[5, :n[5]], r["desc"])
    assert np.array_equal(bits(a[5, :n[5]]), bits(r["angles"]))
    c.close()
    c2.close()


@pytest.mark.parametrize("f,L", [(1.5, 4), (2.0, 3), (1.1, 6)])
def test_full_path_other_scale_factors(V, O, kitti0, f, L):
    """Pyramid geometry, taps and quotas for scale factors other than 1.2 (the integer-2x resize path included)."""
    img = kitti0[:300, :500]
    c = V.Context(V.make_params(nfeatures=700, scaleFactor=f, nlevels=L, max_width=500, max_height=300, keep_side_arrays=1))
    p = O.params(nfeatures=700, scale_factor=f, nlevels=L)
    _compare_full(c, O, img, p, c.max_kp)
    for l in range(L):
        assert np.array_equal(c.get_level(0, l, 500, 300), O.build_level(img, p, l))
    c.close()


def test_sixteen_levels(V, O, kitti0):
    """nlevels == ORB_MAX_LEVELS (16): the prefix of the per-level kept counts in k_describe spans all 32 lanes, so the
    total includes level 0 (a 16-lane scan dropped it: ADVICE r1)."""
    img = kitti0[:300, :700]
    c = V.Context(V.make_params(nfeatures=1500, scaleFactor=1.12, nlevels=16, max_width=700, max_height=300, keep_side_arrays=1))
    p = O.params(nfeatures=1500, scale_factor=1.12, nlevels=16)
    n = _compare_full(c, O, img, p, c.max_kp)
    assert n > 1000
    c.close()


def test_shape_changes_on_one_context(V, O, kitti0):
    """A context re-plans (level geometry, tap and tile tables) whenever the frame shape changes."""
    c = V.Context(V.make_params(nfeatures=800, nlevels=6, max_width=1241, max_height=376, max_batch=3, keep_side_arrays=1))
    p = O.params(nfeatures=800, nlevels=6)
    for shape in ((376, 1241), (200, 640), (376, 1241), (123, 517)):
        img = kitti0[:shape[0], :shape[1]]
        _compare_full(c, O, img, p, c.max_kp)
    frames = np.stack([kitti0[:200, :640], kitti0[100:300, 300:940], kitti0[50:250, 10:650]])
    k, a, d, n = c.detect_and_compute_batch(frames)
    for i in range(3):
        r = O.detect_and_compute(frames[i], p, cap=c.max_kp)
        assert n[i] == r["n"] and np.array_equal(d[i, :n[i]], r["desc"]) and np.array_equal(k[i, :n[i]], r["kps"])
    c.close()
