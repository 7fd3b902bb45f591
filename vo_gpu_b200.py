"""Import alias for the hyphenated package directory ``visual-odometry-gpu_b200``."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("visual-odometry-gpu_b200")
globals().update({k: getattr(_pkg, k) for k in _pkg.__all__})
pkg = _pkg
