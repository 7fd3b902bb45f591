"""Sanitizer substitute (compute-sanitizer is closed on the B200 pool): the parity suite on a -DORB_BOUNDS_CHECK build of the
library, in which every shared / global index the ORB kernels compute (k_pyramid, k_fast, k_edges, k_harris, k_select,
k_describe: tile and window addresses, the byte after a row's last pixel that k_pyramid's second tap may read, cp.async
source ranges, list and table slots, output records) is checked against its buffer before the access.  The suite must
pass with zero failed checks."""
import importlib
import os
import re
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_parity_suite_on_bounds_check_build(tmp_path):
    B = importlib.import_module("visual-odometry-gpu_b200.build")
    lib = B.build(force=True, extra=["-DORB_BOUNDS_CHECK"], out=str(tmp_path / "liborb_b200.so"))
    env = dict(os.environ, ORB_B200_LIB=lib, ORB_EXPECT_BOUNDS_CHECK="1")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(ROOT, "tests", "test_gpu_parity.py"), "-m", "gpu", "-q", "-x", "-s",
                        "-p", "no:cacheprovider"], cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=1500)
    tail = r.stdout[-3000:]
    assert r.returncode == 0, tail
    m = re.search(r"orb bounds check: enabled=(\d+) failures=(\d+) first_line=(\d+) ctas_checked=(\d+)", r.stdout)
    assert m, tail
    enabled, failures, first_line, ctas = (int(v) for v in m.groups())
    assert enabled == 1 and ctas > 10000, (enabled, ctas)
    assert failures == 0, "bounds check failed first at orb_kernels.cuh:%d (%d failures)" % (first_line, failures)
    # the counters do count: a kernel whose check fails on purpose
    code = ("import importlib; V = importlib.import_module('visual-odometry-gpu_b200'); c = V.Context(V.make_params(max_width=64, max_height=64)); "
            "a = c.bounds_check(); c.bounds_selftest(); b = c.bounds_check(); print('selftest', a[1], b[1], b[2])")
    r2 = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=300)
    m2 = re.search(r"selftest (\d+) (\d+) (\d+)", r2.stdout)
    assert m2, r2.stdout[-2000:]
    assert int(m2.group(1)) == 0 and int(m2.group(2)) == 1 and int(m2.group(3)) > 0


def test_normal_build_has_no_checks(V):
    c = V.Context(V.make_params(max_width=64, max_height=64))
    enabled, failures, _, _ = c.bounds_check()
    c.close()
    assert enabled == (1 if os.environ.get("ORB_EXPECT_BOUNDS_CHECK") else 0) and failures == 0
