"""visual-odometry-gpu_b200: B200-native ORB feature extractor (drop-in for the hand-written ORB path of
WeeFav/Visual-Odometry-GPU: include/orb.hpp / include/orb_cpu.hpp).

The directory name carries a hyphen (it is the reference's name); import it with
``importlib.import_module("visual-odometry-gpu_b200")`` or through the alias module ``vo_gpu_b200`` at the
repository root.  The product is csrc/ (CUDA kernels + C ABI, built to liborb_b200.so) and the thin host
mirror in orb.py; the CPU parity checker is never imported from here.
"""
from .orb import (KP, MATCH, ORB, ORBCPU, Context, OrbError, OrientedFAST, Params, RotatedBRIEF, SELECT_HARRIS_TOP_N,
                  SELECT_RASTER_FIRST_N, EXPORTS, default_params, imdecode_gray8, imread_gray8, lib_path, load_library, make_params,
                  png_info)
from .sharding import shard_range
from . import synth
from .synth import synth_frames

__all__ = ["KP", "MATCH", "ORB", "ORBCPU", "Context", "OrbError", "OrientedFAST", "Params", "RotatedBRIEF",
           "SELECT_HARRIS_TOP_N", "SELECT_RASTER_FIRST_N", "EXPORTS", "default_params", "lib_path", "load_library",
           "make_params", "shard_range", "synth_frames", "imdecode_gray8", "imread_gray8", "png_info"]
