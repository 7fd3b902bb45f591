"""The three OpenCV primitives on the path, restated in the oracle, against Python cv2 as witness
(resize INTER_LINEAR, GaussianBlur 5x5 sigma 0, integral).  cv2 is test-only."""
import hashlib

import numpy as np
import pytest

from conftest import noise_image

cv2 = pytest.importorskip("cv2")


def _check_pyramid(O, img, f, L):
    H, W = img.shape
    for l in range(1, L):
        w, h = O.level_size(W, H, f, l)
        a = O.resize_linear(img, w, h)
        b = cv2.resize(img, (w, h), interpolation=cv2.INTER_LINEAR)
        assert np.array_equal(a, b), ("resize", l)
        assert np.array_equal(O.gauss5x5(a), cv2.GaussianBlur(b, (5, 5), 0)), ("blur", l)


def test_resize_blur_kitti(O, kitti0):
    _check_pyramid(O, kitti0, 1.2, 8)


@pytest.mark.parametrize("shape,f,L", [((376, 1241), 1.2, 8), ((370, 1226), 1.2, 8), ((270, 480), 1.2, 8),
                                       ((333, 517), 1.5, 5), ((256, 320), 2.0, 4), ((97, 131), 1.2, 12)])
def test_resize_blur_noise(O, shape, f, L):
    _check_pyramid(O, noise_image(shape[0], shape[1], 11), f, L)


def test_blur_small_images(O):
    for (h, w) in [(5, 5), (3, 7), (2, 9), (1, 6), (6, 1), (17, 33)]:
        a = noise_image(h, w, h * 100 + w)
        assert np.array_equal(O.gauss5x5(a), cv2.GaussianBlur(a, (5, 5), 0))


def test_integral(O, kitti0):
    ii = O.integral_flat(kitti0)
    assert np.array_equal(ii[:377], cv2.integral(kitti0)) and not ii[377:].any()


def test_golden_pyramid_hashes(O, golden, kitti0, kitti1):
    """cv2's pyramid, frozen as sha256 in the golden file, equals the oracle's (works without cv2 resize too)."""
    for tag, img in (("k0", kitti0), ("k1", kitti1)):
        p = O.params(nlevels=8)
        for l in range(1, 8):
            lv = O.build_level(img, p, l)
            assert hashlib.sha256(lv.tobytes()).hexdigest() == str(golden["cv2_%s_pyr_sha" % tag][l - 1])
    img = np.random.default_rng(7).integers(0, 256, (1080, 1920), dtype=np.uint8)
    p = O.params(nlevels=8)
    for l in (1, 4, 7):
        assert hashlib.sha256(O.build_level(img, p, l).tobytes()).hexdigest() == str(golden["cv2_noise1080_pyr_sha"][l - 1])


def test_libm_twins_host_side(O):
    """lround trick used on the device (v + copysign(pred(0.5)) truncated) equals std::lround on a float sweep."""
    rng = np.random.default_rng(3)
    v = np.concatenate([rng.uniform(-40, 40, 200000).astype(np.float32),
                        (np.arange(-80, 81) / 2).astype(np.float32),
                        np.nextafter(np.float32(0.5), np.float32(0)).reshape(1),
                        -np.nextafter(np.float32(0.5), np.float32(0)).reshape(1)])
    h = np.where(np.signbit(v), -np.float32(0.49999997), np.float32(0.49999997)).astype(np.float32)
    trick = np.trunc((v + h).astype(np.float32)).astype(np.int64)
    ref = np.array([O.lib().orc_lround_f(float(x)) for x in v[:4000]])
    assert np.array_equal(trick[:4000], ref)
    half_away = np.where(v >= 0, np.floor(v.astype(np.float64) + 0.5), np.ceil(v.astype(np.float64) - 0.5)).astype(np.int64)
    assert np.array_equal(trick, half_away)
