#!/usr/bin/env python3
"""Generate tests/golden/golden_v1.npz (run in the build container, where /root/reference and cv2 exist).

Sources of truth recorded in the file:
  ref_*   : outputs of the reference's OWN src/orb_cpu.cpp, compiled unmodified into oracle/_ref
            (ORBCPU().detectAndCompute = single level, thr 50, cap 3000, patch 9; and the stage classes at
            thr 20 / patch 31) on 000000.png and 000001.png;
  cv2_*   : sha256 of cv2.resize(INTER_LINEAR) -> cv2.GaussianBlur(5x5, 0) for every pyramid level (cv2 4.13);
  ml_*    : outputs of the multi-level oracle composition for both parameter sets.
The GPU tests compare the CUDA path against these files on the B200 box, where /root/reference is absent.
"""
import hashlib
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as O  # noqa: E402


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    out = {}
    assert O.have_ref(), "needs /root/reference to build oracle/_ref"
    for tag, name in (("k0", "000000.png"), ("k1", "000001.png")):
        img = cv2.imread("/root/reference/" + name, cv2.IMREAD_GRAYSCALE)
        assert sha(img) == sha(cv2.imread(os.path.join(ROOT, "tests/golden/kitti_" + name), cv2.IMREAD_GRAYSCALE))
        # the reference itself, as shipped
        k, a, d = O.ref_orbcpu(img)
        out["ref_%s_t50_kps" % tag], out["ref_%s_t50_ang" % tag], out["ref_%s_t50_desc" % tag] = k, a, d
        # the reference's stage classes at the orb.hpp defaults (thr 20, patch 31), cap 3000
        k = O.ref_fast_detect(img, 3000, 20)
        a = O.ref_orientations(img, k, 31)
        d = O.ref_brief(img, k, a)
        out["ref_%s_t20_kps" % tag], out["ref_%s_t20_ang" % tag], out["ref_%s_t20_desc" % tag] = k, a, d
        # cv2 witness of the pyramid
        hs = []
        for l in range(1, 8):
            w, h = O.level_size(img.shape[1], img.shape[0], 1.2, l)
            lv = cv2.GaussianBlur(cv2.resize(img, (w, h), interpolation=cv2.INTER_LINEAR), (5, 5), 0)
            hs.append(sha(lv))
        out["cv2_%s_pyr_sha" % tag] = np.array(hs)
        # multi-level oracle, both parameter sets (D2), N = 2000, 8 levels, Harris top-N
        for pt, thr, patch in (("t20", 20, 31), ("t50", 50, 9)):
            p = O.params(nfeatures=2000, nlevels=8, fast_threshold=thr, orient_patch=patch, select_policy=1)
            r = O.detect_and_compute(img, p, cap=2000)
            for key in ("kps", "angles", "desc", "n_per_level", "level_xy", "level_id", "response"):
                out["ml_%s_%s_%s" % (tag, pt, key)] = r[key]
    # a seeded 1080p noise frame: pyramid witness only (cv2)
    rng = np.random.default_rng(7)
    img = rng.integers(0, 256, (1080, 1920), dtype=np.uint8)
    hs = []
    for l in range(1, 8):
        w, h = O.level_size(1920, 1080, 1.2, l)
        hs.append(sha(cv2.GaussianBlur(cv2.resize(img, (w, h), interpolation=cv2.INTER_LINEAR), (5, 5), 0)))
    out["cv2_noise1080_pyr_sha"] = np.array(hs)
    path = os.path.join(ROOT, "tests/golden/golden_v1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes;", len(out), "arrays")


if __name__ == "__main__":
    main()
