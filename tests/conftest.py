import importlib
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session", autouse=True)
def _bounds_check_report():
    """On a -DORB_BOUNDS_CHECK build (tests/test_gpu_bounds.py sets ORB_EXPECT_BOUNDS_CHECK) print the device-side counters after
    the last test of the session."""
    yield
    if os.environ.get("ORB_EXPECT_BOUNDS_CHECK"):
        Vm = importlib.import_module("visual-odometry-gpu_b200")
        c = Vm.Context(Vm.make_params(max_width=64, max_height=64))
        print("\norb bounds check: enabled=%d failures=%d first_line=%d ctas_checked=%d" % c.bounds_check())
        c.close()


@pytest.fixture(scope="session")
def O():
    """The CPU parity oracle (test infrastructure; oracle/pyoracle.py)."""
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def V():
    """The product package (hyphenated directory name -> importlib)."""
    return importlib.import_module("visual-odometry-gpu_b200")


def _png(name):
    import cv2
    img = cv2.imread(os.path.join(GOLDEN, name), cv2.IMREAD_GRAYSCALE)
    assert img is not None and img.dtype == np.uint8
    return img


@pytest.fixture(scope="session")
def kitti0():
    return _png("kitti_000000.png")


@pytest.fixture(scope="session")
def kitti1():
    return _png("kitti_000001.png")


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(GOLDEN, "golden_v1.npz"))


def noise_image(h, w, seed, kind="uniform"):
    rng = np.random.default_rng(seed)
    if kind == "uniform":
        return rng.integers(0, 256, (h, w), dtype=np.uint8)
    if kind == "blocks":   # flat rectangles: sparse, strong corners, many exact score ties
        img = np.full((h, w), 90, np.uint8)
        for _ in range(max(4, h * w // 4000)):
            y, x = int(rng.integers(0, h)), int(rng.integers(0, w))
            hh, ww = int(rng.integers(4, 40)), int(rng.integers(4, 40))
            img[y:y + hh, x:x + ww] = rng.integers(0, 256)
        return img
    raise ValueError(kind)
