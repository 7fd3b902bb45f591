"""Pin the oracle: (1) against the reference's own src/orb_cpu.cpp compiled unmodified (oracle/_ref, only where
it was built), (2) against the committed golden fixtures generated from it, (3) against the survey's known-answer
values."""
import hashlib

import numpy as np
import pytest

from conftest import noise_image


def test_pattern_table(O):
    pat = O.pattern()
    assert pat.shape == (1024,) and pat.min() == -13 and pat.max() == 12
    h = hashlib.sha256(",".join(str(int(v)) for v in pat).encode()).hexdigest()
    assert h == "88df8ca875cc8db56799edd57bb914edad8acb2d48c202b7a464a575b55dbdb8"
    if O.have_ref():
        assert np.array_equal(pat.astype(np.int32), O.ref_pattern())


def test_geometry_matches_survey(O):
    # level sizes and quotas
    assert [O.level_size(1241, 376, 1.2, l) for l in range(8)] == [
        (1241, 376), (1034, 313), (862, 261), (718, 218), (598, 181), (499, 151), (416, 126), (346, 105)]
    assert [O.level_size(1226, 370, 1.2, l) for l in range(8)] == [
        (1226, 370), (1022, 308), (851, 257), (709, 214), (591, 178), (493, 149), (411, 124), (342, 103)]
    assert [O.level_size(3840, 2160, 1.2, l)[0] for l in range(12)] == [
        3840, 3200, 2667, 2222, 1852, 1543, 1286, 1072, 893, 744, 620, 517]
    assert [O.level_quota(2000, 1.2, 8, l) for l in range(8)] == [434, 361, 301, 251, 209, 174, 145, 121]
    assert [O.level_quota(5000, 1.2, 8, l) for l in range(8)] == [1085, 904, 754, 628, 523, 436, 363, 303]
    assert [O.level_quota(10000, 1.2, 12, l) for l in range(12)] == [
        1877, 1564, 1303, 1086, 905, 754, 628, 523, 436, 363, 303, 252]


@pytest.mark.parametrize("tag", ["k0", "k1"])
def test_single_level_equals_reference_golden(O, golden, kitti0, kitti1, tag):
    """D3: nlevels=1, raster-first 3000, thr 50, patch 9 == ORBCPU::detectAndCompute as shipped."""
    img = kitti0 if tag == "k0" else kitti1
    p = O.params(nfeatures=3000, nlevels=1, fast_threshold=50, orient_patch=9, select_policy=0)
    r = O.detect_and_compute(img, p, cap=3000)
    assert np.array_equal(r["kps"], golden["ref_%s_t50_kps" % tag])
    assert np.array_equal(r["angles"].view(np.uint32), golden["ref_%s_t50_ang" % tag].view(np.uint32))
    assert np.array_equal(r["desc"], golden["ref_%s_t50_desc" % tag])
    p = O.params(nfeatures=3000, nlevels=1, fast_threshold=20, orient_patch=31, select_policy=0)
    r = O.detect_and_compute(img, p, cap=3000)
    assert np.array_equal(r["kps"], golden["ref_%s_t20_kps" % tag])
    assert np.array_equal(r["angles"].view(np.uint32), golden["ref_%s_t20_ang" % tag].view(np.uint32))
    assert np.array_equal(r["desc"], golden["ref_%s_t20_desc" % tag])


def test_survey_known_answers(O, kitti0, kitti1):
    """Counts from the survey's independent transliteration (SURVEY.md 8(c) known-answer row)."""
    p = O.params(nfeatures=3000, nlevels=1, fast_threshold=50, orient_patch=9, select_policy=0)
    r = O.detect_and_compute(kitti0, p, cap=3000)
    assert r["n"] == 1178 and tuple(r["kps"][0]) == (815, 3) and r["angles"][0] == 0.0
    assert r["desc"][0, :8].tobytes().hex() == "0088120022480084"
    assert int(np.unpackbits(r["desc"]).sum()) == 146595
    fl = O.brief_flags(1241, 376, r["kps"], r["angles"])
    assert int((fl & 1).astype(bool).sum()) == 64 and int((fl & 2).astype(bool).sum()) == 6
    r = O.detect_and_compute(kitti1, p, cap=3000)
    assert r["n"] == 1233 and tuple(r["kps"][0]) == (1232, 3)
    assert int(np.unpackbits(r["desc"]).sum()) == 151981
    p = O.params(nfeatures=3000, nlevels=1, fast_threshold=20, orient_patch=31, select_policy=0)
    r = O.detect_and_compute(kitti0, p, cap=3000)
    assert r["n"] == 3000 and tuple(r["kps"][-1]) == (394, 174)
    assert int(np.unpackbits(r["desc"]).sum()) == 408672
    s0, s1 = O.fast_scores(kitti0, 20), O.fast_scores(kitti1, 20)
    assert int((s0 > 0).sum()) == 14415 and int((s1 > 0).sum()) == 14587
    assert len(O.nms(s0)) == 4153 and len(O.nms(s1)) == 4227
    assert int((O.fast_scores(kitti0, 50) > 0).sum()) == 3855 and int((O.fast_scores(kitti1, 50) > 0).sum()) == 3863


def test_multilevel_matches_golden(O, golden, kitti0):
    for pt, thr, patch in (("t20", 20, 31), ("t50", 50, 9)):
        p = O.params(nfeatures=2000, nlevels=8, fast_threshold=thr, orient_patch=patch, select_policy=1)
        r = O.detect_and_compute(kitti0, p, cap=2000)
        for key in ("kps", "desc", "n_per_level", "level_xy", "level_id"):
            assert np.array_equal(r[key], golden["ml_k0_%s_%s" % (pt, key)]), key
        assert np.array_equal(r["angles"].view(np.uint32), golden["ml_k0_%s_angles" % pt].view(np.uint32))
        assert np.array_equal(r["response"].view(np.uint32), golden["ml_k0_%s_response" % pt].view(np.uint32))
    # thr 20 fills every level's quota on the real frame (SURVEY.md 8 candidate densities)
    assert list(golden["ml_k0_t20_n_per_level"]) == [434, 361, 301, 251, 209, 174, 145, 121]


@pytest.mark.skipif("not __import__('oracle.pyoracle').pyoracle.have_ref()")
def test_oracle_equals_compiled_reference_on_random_frames(O):
    """Only where oracle/_ref exists: stage-by-stage equality with the reference's classes on synthetic input."""
    for seed, kind, thr in ((1, "uniform", 50), (2, "blocks", 20), (3, "uniform", 20)):
        img = noise_image(120, 173, seed, kind)
        k_ref = O.ref_fast_detect(img, 3000, thr)
        k_or = O.nms(O.fast_scores(img, thr), 3, 3000)
        assert np.array_equal(k_ref, k_or)
        for patch in (9, 31):
            a_ref, a_or = O.ref_orientations(img, k_ref, patch), O.orientations(img, k_ref, patch)
            assert np.array_equal(a_ref.view(np.uint32), a_or.view(np.uint32))
        assert np.array_equal(O.ref_brief(img, k_ref, a_ref), O.brief(img, k_ref, a_or))


def test_selection_properties(O, kitti0):
    p = O.params(nfeatures=600, nlevels=4, fast_threshold=20, orient_patch=31, select_policy=1)
    r = O.detect_and_compute(kitti0, p, cap=600)
    lid = r["level_id"]
    assert np.all(np.diff(lid) >= 0)                                   # levels concatenated 0..L-1
    for l in range(4):
        xy = r["level_xy"][lid == l]
        key = xy["y"].astype(np.int64) * 65536 + xy["x"]
        assert np.all(np.diff(key) > 0)                                # raster order inside a level
        assert len(xy) == min(O.level_quota(600, 1.2, 4, l), len(O.nms(O.fast_scores(O.build_level(kitti0, p, l), 20))))
        sc = O.level_scale(1.2, l)
        k = r["kps"][lid == l]
        assert np.array_equal(k["x"], (xy["x"].astype(np.float32) * np.float32(sc)).astype(np.int32))   # src/orb.cpp:96


def test_degenerate_inputs(O):
    p = O.params(nfeatures=100, nlevels=3, fast_threshold=20, orient_patch=31, select_policy=1)
    flat = np.full((40, 50), 128, np.uint8)
    assert O.detect_and_compute(flat, p, cap=100)["n"] == 0             # no corners
    tiny = noise_image(9, 11, 5)
    r = O.detect_and_compute(tiny, p, cap=100)                          # levels shrink below the FAST border
    assert r["n"] >= 0
    img = noise_image(64, 64, 6)
    r = O.detect_and_compute(img, O.params(nfeatures=1000, nlevels=1, select_policy=0), cap=10)
    assert r["n"] == 10                                                 # output cap truncates
