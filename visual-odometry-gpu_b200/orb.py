"""Host-side mirror of the reference's ORB interface over the C ABI (include/orb_b200.h).

Class and method names follow the reference's C++ facade (include/orb.hpp:10-49 and the CPU twins
include/orb_cpu.hpp:4-43): ``ORB(nfeatures, scaleFactor, nlevels).detectAndCompute(image)``,
``OrientedFAST(threshold, n, nms_window, patch_size).detect(image, nfeatures)`` /
``.compute_orientations(image, keypoints)``, ``RotatedBRIEF().compute(image, keypoints, orientations)``.
Images are numpy uint8 2-D arrays (the cv::Mat CV_8UC1 of the reference); keypoints come back as a
structured array with fields x, y (== Keypoint, include/orb.hpp:4), descriptors as (N, 32) uint8
(== ORBDescriptor, include/orb.hpp:6-8).

There is no CPU fallback: everything below calls liborb_b200.so, and creating a context without a
CUDA device raises OrbError.  The C++ facade with the reference's exact signatures is
include/orb.hpp of this repository; both sit on the same C ABI.
"""
import ctypes as C
import os

import numpy as np

from . import build as _build

KP = np.dtype([("x", "<i4"), ("y", "<i4")])
MATCH = np.dtype([("idx1", "<i4"), ("dist1", "<i4"), ("idx2", "<i4"), ("dist2", "<i4")])   # == orb_match

SELECT_RASTER_FIRST_N = 0
SELECT_HARRIS_TOP_N = 1


class OrbError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("orb_b200 error %d: %s" % (code, msg))
        self.code = code


class Params(C.Structure):
    _fields_ = [("nfeatures", C.c_int32), ("scale_factor", C.c_float), ("nlevels", C.c_int32),
                ("fast_threshold", C.c_int32), ("fast_n", C.c_int32), ("nms_window", C.c_int32),
                ("orient_patch", C.c_int32), ("select_policy", C.c_int32), ("blur_levels", C.c_int32),
                ("harris_k", C.c_float), ("device", C.c_int32), ("max_width", C.c_int32),
                ("max_height", C.c_int32), ("max_batch", C.c_int32), ("chunk_frames", C.c_int32),
                ("max_keypoints", C.c_int32), ("keep_side_arrays", C.c_int32), ("reserved", C.c_int32 * 3)]


class ImageInfo(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("bit_depth", C.c_int32), ("channels", C.c_int32)]


EXPORTS = [
    "orb_abi_version", "orb_default_params", "orb_create", "orb_destroy", "orb_last_error", "orb_set_stream", "orb_use_own_stream",
    "orb_synchronize", "orb_detect_and_compute", "orb_detect_and_compute_batch", "orb_get_level", "orb_level_size",
    "orb_level_quota", "orb_fast_detect", "orb_nms_scores", "orb_conv2d_u8", "orb_gaussian_blur_1d", "orb_harris", "orb_orientations", "orb_brief", "orb_get_side_arrays",
    "orb_get_candidates", "orb_get_harris_weights", "orb_last_launch_count", "orb_set_profiling", "orb_get_stage_ms", "orb_match_knn2", "orb_match_knn2_batch", "orb_ratio_test", "orb_debug_eval_math", "orb_debug_bounds_check", "orb_debug_bounds_selftest", "bit_pattern_31_",
    "orb_png_info", "orb_png_decode_gray8", "orb_imread_gray8", "orb_detect_and_compute_files", "orb_get_ingested_frame", "orb_debug_inflate", "orb_lk_track", "orb_lk_track_batch", "orb_lk_levels", "orb_lk_get_level", "orb_debug_wave_schedule",
]

_lib = None


def lib_path():
    """The library in use: in-tree liborb_b200.so, or the variant build named by ORB_B200_LIB (tests/test_gpu_bounds.py)."""
    return os.environ.get("ORB_B200_LIB") or _build.LIB


def load_library():
    """dlopen liborb_b200.so (building it in-tree with nvcc if it is missing or stale)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("ORB_B200_LIB") or _build.build()
    L = C.CDLL(path)
    vp, i, sz = C.c_void_p, C.c_int, C.c_size_t
    L.orb_abi_version.restype = i
    L.orb_default_params.argtypes = [C.POINTER(Params)]
    L.orb_default_params.restype = None
    L.orb_create.argtypes = [C.POINTER(Params), C.POINTER(vp)]
    L.orb_destroy.argtypes = [vp]
    L.orb_destroy.restype = None
    L.orb_last_error.argtypes = [vp]
    L.orb_last_error.restype = C.c_char_p
    L.orb_set_stream.argtypes = [vp, vp]
    L.orb_use_own_stream.argtypes = [vp]
    L.orb_synchronize.argtypes = [vp]
    L.orb_detect_and_compute.argtypes = [vp, vp, i, i, sz, i, vp, vp, vp, vp, vp]
    L.orb_detect_and_compute_batch.argtypes = [vp, vp, i, i, i, i, sz, sz, i, vp, vp, vp, vp, i]
    L.orb_get_level.argtypes = [vp, i, i, vp, sz, C.POINTER(i), C.POINTER(i)]
    L.orb_level_size.argtypes = [vp, i, i, i, C.POINTER(i), C.POINTER(i)]
    L.orb_level_quota.argtypes = [vp, i]
    L.orb_fast_detect.argtypes = [vp, vp, i, i, sz, i, vp, C.POINTER(i)]
    L.orb_harris.argtypes = [vp, vp, i, i, sz, vp, i, vp]
    L.orb_orientations.argtypes = [vp, vp, i, i, sz, vp, i, vp]
    L.orb_brief.argtypes = [vp, vp, i, i, sz, vp, vp, i, vp]
    L.orb_get_side_arrays.argtypes = [vp, i, i, vp, vp, vp]
    L.orb_get_candidates.argtypes = [vp, i, i, i, vp, vp, C.POINTER(i)]
    L.orb_get_harris_weights.argtypes = [vp, vp]
    L.orb_last_launch_count.argtypes = [vp]
    L.orb_debug_eval_math.argtypes = [vp, i, vp, vp, i, vp]
    L.orb_debug_bounds_check.argtypes = [vp, vp, vp, vp, vp]
    L.orb_debug_bounds_selftest.argtypes = [vp]
    L.orb_match_knn2.argtypes = [vp, vp, i, vp, i, i, vp]
    L.orb_match_knn2_batch.argtypes = [vp, vp, vp, i, i, i, vp]
    L.orb_ratio_test.argtypes = [vp, i, C.c_float, vp]
    L.orb_ratio_test.restype = None
    L.orb_set_profiling.argtypes = [vp, i]
    L.orb_get_stage_ms.argtypes = [vp, C.POINTER(C.c_float * 5), C.POINTER(C.c_int * 5)]
    L.orb_png_info.argtypes = [vp, sz, C.POINTER(ImageInfo)]
    L.orb_png_decode_gray8.argtypes = [vp, sz, vp, sz, i, i]
    L.orb_imread_gray8.argtypes = [C.c_char_p, vp, sz, i, i, C.POINTER(i), C.POINTER(i)]
    L.orb_detect_and_compute_files.argtypes = [vp, C.POINTER(C.c_char_p), i, i, i, i, vp, vp, vp, vp, i]
    L.orb_get_ingested_frame.argtypes = [vp, i, vp, sz, C.POINTER(i), C.POINTER(i)]
    L.orb_debug_inflate.argtypes = [vp, vp, vp, i, vp, vp, vp]
    L.orb_lk_track.argtypes = [vp, vp, vp, i, i, sz, vp, i, i, i, i, C.c_double, C.c_float, vp, vp, vp]
    L.orb_debug_wave_schedule.argtypes = [i, i, i, i, vp, i]
    L.orb_lk_track_batch.argtypes = [vp, vp, i, i, i, i, sz, sz, vp, vp, i, i, i, i, i, C.c_double, C.c_float, vp, vp, vp]
    L.orb_lk_levels.argtypes = [i, i, i, i]
    L.orb_lk_get_level.argtypes = [vp, i, i, vp, C.POINTER(i), C.POINTER(i)]
    _lib = L
    return L


def wave_schedule(n_frames, wave, ramp_up, ramp_down):
    """Wave boundaries the batch pipeline uses (host logic, no GPU needed): list of begins ending with n_frames."""
    lib = load_library()
    buf = (C.c_int * (n_frames + 2))()
    n = lib.orb_debug_wave_schedule(n_frames, wave, int(ramp_up), int(ramp_down), buf, n_frames + 2)
    return list(buf[:n])


def default_params():
    p = Params()
    load_library().orb_default_params(C.byref(p))
    return p


def _ck_global(rc):
    if rc != 0:
        raise OrbError(rc, load_library().orb_last_error(None).decode())


def png_info(data):
    """(width, height, bit_depth, channels) of an encoded PNG held in a bytes object."""
    info = ImageInfo()
    _ck_global(load_library().orb_png_info(data, len(data), C.byref(info)))
    return info.width, info.height, info.bit_depth, info.channels


def imdecode_gray8(data):
    """== cv2.imdecode(data, cv2.IMREAD_GRAYSCALE) for the PNG layouts of include/orb_b200.h (host decode)."""
    w, h, _, _ = png_info(data)
    out = np.empty((h, w), np.uint8)
    _ck_global(load_library().orb_png_decode_gray8(data, len(data), _p(out), out.strides[0], w, h))
    return out


def imread_gray8(path):
    """== cv::imread(path, cv::IMREAD_GRAYSCALE), reference src/feature_matching.cpp:55 (host decode)."""
    with open(path, "rb") as f:
        head = f.read(64)
    w, h, _, _ = png_info(head)
    out = np.empty((h, w), np.uint8)
    ww, hh = C.c_int(), C.c_int()
    _ck_global(load_library().orb_imread_gray8(os.fsencode(path), _p(out), out.strides[0], w, h, C.byref(ww), C.byref(hh)))
    return out


def _img(a):
    a = np.asarray(a)
    if a.dtype != np.uint8 or a.ndim != 2:
        raise ValueError("image must be a 2-D uint8 array (CV_8UC1)")   # CV_Assert at ref src/orb_cpu.cpp:26
    if a.strides[1] != 1:
        a = np.ascontiguousarray(a)
    return a


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Context:
    """One orb_ctx: device arena + stream + plan.  Not thread-safe; one per (device, host thread)."""

    def __init__(self, params):
        self.lib = load_library()
        self.params = params
        h = C.c_void_p()
        rc = self.lib.orb_create(C.byref(params), C.byref(h))
        if rc != 0:
            raise OrbError(rc, self.lib.orb_last_error(None).decode())
        self.h = h
        tq = sum(self.level_quota(l) for l in range(params.nlevels))
        self.max_kp = params.max_keypoints if params.max_keypoints > 0 else min(tq, 8192 * params.nlevels)

    def close(self):
        if getattr(self, "h", None):
            self.lib.orb_destroy(self.h)
            self.h = None

    __del__ = close

    def _ck(self, rc):
        if rc != 0:
            raise OrbError(rc, self.lib.orb_last_error(self.h).decode())

    def set_stream(self, cuda_stream):
        self._ck(self.lib.orb_set_stream(self.h, C.c_void_p(cuda_stream)))

    def use_own_stream(self):
        self._ck(self.lib.orb_use_own_stream(self.h))

    def synchronize(self):
        self._ck(self.lib.orb_synchronize(self.h))

    def level_size(self, w, h, level):
        lw, lh = C.c_int(), C.c_int()
        self._ck(self.lib.orb_level_size(self.h, w, h, level, C.byref(lw), C.byref(lh)))
        return lw.value, lh.value

    def level_quota(self, level):
        return self.lib.orb_level_quota(self.h, level)

    def launch_count(self):
        return self.lib.orb_last_launch_count(self.h)

    def set_profiling(self, on):
        self._ck(self.lib.orb_set_profiling(self.h, int(on)))

    def stage_ms(self):
        """(ms[5], launches[5]) per kernel (pyramid, FAST, Harris, select, describe) since the last call; needs set_profiling(True)."""
        ms, n = (C.c_float * 5)(), (C.c_int * 5)()
        self._ck(self.lib.orb_get_stage_ms(self.h, C.byref(ms), C.byref(n)))
        return list(ms), list(n)

    # ---- whole path -----------------------------------------------------------------------
    def detect_and_compute(self, image, cap=None):
        img = _img(image)
        cap = cap or self.max_kp
        kps = np.zeros(cap, KP)
        ang = np.zeros(cap, np.float32)
        des = np.zeros((cap, 32), np.uint8)
        n = C.c_int()
        npl = np.zeros(self.params.nlevels, np.int32)
        self._ck(self.lib.orb_detect_and_compute(self.h, _p(img), img.shape[1], img.shape[0], img.strides[0], cap,
                                                 _p(kps), _p(ang), _p(des), C.byref(n), _p(npl)))
        return kps[:n.value], ang[:n.value], des[:n.value], npl

    def detect_and_compute_batch(self, frames, cap=None, out=None):
        """frames: (F, H, W) uint8 host array.  Returns (kps[F,cap], angles[F,cap], desc[F,cap,32], n[F])."""
        frames = np.asarray(frames)
        if frames.dtype != np.uint8 or frames.ndim != 3 or frames.strides[2] != 1:
            raise ValueError("frames must be a (F, H, W) uint8 array")
        F, H, W = frames.shape
        cap = cap or self.max_kp
        if out is None:
            out = (np.zeros((F, cap), KP), np.zeros((F, cap), np.float32), np.zeros((F, cap, 32), np.uint8),
                   np.zeros(F, np.int32))
        kps, ang, des, n = out
        self._ck(self.lib.orb_detect_and_compute_batch(self.h, _p(frames), 0, F, W, H, frames.strides[1],
                                                       frames.strides[0], cap, _p(kps), _p(ang), _p(des), _p(n), 0))
        return kps, ang, des, n

    def detect_and_compute_files(self, paths, cap=None, threads=0, decode_on_device=False, out=None):
        """PNG files of one size -> (kps[F,cap], angles[F,cap], desc[F,cap,32], n[F]); decode overlaps the device work."""
        F = len(paths)
        cap = cap or self.max_kp
        arr = (C.c_char_p * F)(*[os.fsencode(p) for p in paths])
        if out is None:
            out = (np.zeros((F, cap), KP), np.zeros((F, cap), np.float32), np.zeros((F, cap, 32), np.uint8),
                   np.zeros(F, np.int32))
        kps, ang, des, n = out
        self._ck(self.lib.orb_detect_and_compute_files(self.h, arr, F, int(threads), int(bool(decode_on_device)), cap,
                                                       _p(kps), _p(ang), _p(des), _p(n), 0))
        return kps, ang, des, n

    def lk_track(self, prev, nxt, pts, win=21, max_level=3, max_iter=30, eps=0.01, min_eig=0.001):
        """== cv2.calcOpticalFlowPyrLK(prev, nxt, pts, None, winSize=(win, win), maxLevel=max_level,
        criteria=(COUNT+EPS, max_iter, eps), flags=0, minEigThreshold=min_eig) (reference src/feature_tracking.cpp:174-180):
        returns (next_pts [n,2] float32, status [n] uint8, err [n] float32)."""
        prev, nxt = _img(prev), _img(nxt)
        if prev.shape != nxt.shape or prev.strides[0] != nxt.strides[0]:
            nxt = np.ascontiguousarray(nxt)
            prev = np.ascontiguousarray(prev)
        pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
        out = np.zeros_like(pts)
        st = np.zeros(len(pts), np.uint8)
        er = np.zeros(len(pts), np.float32)
        self._ck(self.lib.orb_lk_track(self.h, _p(prev), _p(nxt), prev.shape[1], prev.shape[0], prev.strides[0], _p(pts),
                                       len(pts), win, max_level, max_iter, eps, min_eig, _p(out), _p(st), _p(er)))
        return out, st, er

    def lk_track_batch(self, frames, pts, n_pts=None, win=21, max_level=3, max_iter=30, eps=0.01, min_eig=0.001):
        """Frame t tracked into frame t + 1 for a whole sequence (host arrays): frames [F, h, w] uint8, pts [F-1, cap, 2] float32,
        n_pts [F-1] (None: cap points everywhere).  Returns (next_pts [F-1, cap, 2], status [F-1, cap], err [F-1, cap])."""
        frames = np.ascontiguousarray(frames, np.uint8)
        F, h, w = frames.shape
        pts = np.ascontiguousarray(pts, np.float32)
        cap = pts.shape[1]
        assert pts.shape == (F - 1, cap, 2)
        n = None if n_pts is None else np.ascontiguousarray(n_pts, np.int32)
        out = np.zeros_like(pts)
        st = np.zeros((F - 1, cap), np.uint8)
        er = np.zeros((F - 1, cap), np.float32)
        self._ck(self.lib.orb_lk_track_batch(self.h, _p(frames), 0, F, w, h, frames.strides[1], frames.strides[0], _p(pts),
                                             None if n is None else _p(n), cap, 0, win, max_level, max_iter, eps, min_eig,
                                             _p(out), _p(st), _p(er)))
        return out, st, er

    def lk_track_batch_ptr(self, frames_ptr, n_frames, w, h, pitch, frame_stride, pts_ptr, n_ptr, cap, next_ptr, status_ptr, err_ptr,
                           win=21, max_level=3, max_iter=30, eps=0.01, min_eig=0.001):
        """device-resident form (asynchronous on the context's stream): frames as given to detect_and_compute_batch_ptr,
        points / counts / results in device memory"""
        self._ck(self.lib.orb_lk_track_batch(self.h, C.c_void_p(frames_ptr), 1, n_frames, w, h, pitch, frame_stride, C.c_void_p(pts_ptr),
                                             C.c_void_p(n_ptr) if n_ptr else None, cap, 1, win, max_level, max_iter, eps, min_eig,
                                             C.c_void_p(next_ptr), C.c_void_p(status_ptr), C.c_void_p(err_ptr) if err_ptr else None))

    def lk_get_level(self, which, level, w0, h0):
        """Pyramid level of the last lk_track call (which: 0 prev, 1 next); w0, h0 = frame size of that call."""
        w, h = C.c_int(), C.c_int()
        tmp = np.zeros(w0 * h0, np.uint8)
        self._ck(self.lib.orb_lk_get_level(self.h, which, level, _p(tmp), C.byref(w), C.byref(h)))
        return tmp[:w.value * h.value].reshape(h.value, w.value).copy()

    def debug_inflate(self, streams, out_sizes):
        """Device inflate of raw deflate streams (test hook): returns (list of bytes, status array)."""
        n = len(streams)
        offs = np.zeros(n + 1, np.uint32)
        offs[1:] = np.cumsum([len(s) for s in streams])
        oofs = np.zeros(n + 1, np.uint32)
        oofs[1:] = np.cumsum(out_sizes)
        blob = np.frombuffer(b"".join(streams) + b"\0", np.uint8)
        out = np.zeros(int(oofs[-1]) + 1, np.uint8)
        st = np.zeros(n, np.int32)
        self._ck(self.lib.orb_debug_inflate(self.h, _p(blob), _p(offs), n, _p(out), _p(oofs), _p(st)))
        return [out[oofs[k]:oofs[k + 1]].tobytes() for k in range(n)], st

    def get_ingested_frame(self, frame, w, h):
        out = np.empty((h, w), np.uint8)
        ww, hh = C.c_int(), C.c_int()
        self._ck(self.lib.orb_get_ingested_frame(self.h, frame, _p(out), out.strides[0], C.byref(ww), C.byref(hh)))
        return out[:hh.value, :ww.value]

    def detect_and_compute_batch_ptr(self, frames_ptr, frames_on_device, n_frames, w, h, pitch, frame_stride, cap,
                                     kps_ptr, ang_ptr, des_ptr, n_ptr, outputs_on_device):
        """Raw-pointer form (device or pinned-host buffers owned by the caller, e.g. torch tensors)."""
        self._ck(self.lib.orb_detect_and_compute_batch(self.h, C.c_void_p(frames_ptr), int(frames_on_device), n_frames,
                                                       w, h, pitch, frame_stride, cap, C.c_void_p(kps_ptr),
                                                       C.c_void_p(ang_ptr), C.c_void_p(des_ptr), C.c_void_p(n_ptr),
                                                       int(outputs_on_device)))

    # ---- read-backs of the last call --------------------------------------------------------
    def get_level(self, frame, level, w, h):
        lw, lh = self.level_size(w, h, level)
        out = np.zeros((lh, lw), np.uint8)
        ow, oh = C.c_int(), C.c_int()
        self._ck(self.lib.orb_get_level(self.h, frame, level, _p(out), lw, C.byref(ow), C.byref(oh)))
        assert (ow.value, oh.value) == (lw, lh)
        return out

    def get_side_arrays(self, frame, n):
        xy = np.zeros(n, KP)
        lid = np.zeros(n, np.int32)
        rsp = np.zeros(n, np.float32)
        self._ck(self.lib.orb_get_side_arrays(self.h, frame, n, _p(xy), _p(lid), _p(rsp)))
        return xy, lid, rsp

    def get_candidates(self, frame, level, cap=1 << 20):
        xy = np.zeros(cap, KP)
        rsp = np.zeros(cap, np.float32)
        n = C.c_int()
        self._ck(self.lib.orb_get_candidates(self.h, frame, level, cap, _p(xy), _p(rsp), C.byref(n)))
        m = min(n.value, cap)
        return xy[:m], rsp[:m], n.value

    def harris_weights(self):
        w = np.zeros(49, np.float32)
        self._ck(self.lib.orb_get_harris_weights(self.h, _p(w)))
        return w

    # ---- stages -------------------------------------------------------------------------------
    def fast_detect(self, image, nfeatures):
        img = _img(image)
        kps = np.zeros(max(nfeatures, 1), KP)
        n = C.c_int()
        self._ck(self.lib.orb_fast_detect(self.h, _p(img), img.shape[1], img.shape[0], img.strides[0], nfeatures,
                                          _p(kps), C.byref(n)))
        return kps[:n.value].copy()

    def nms_scores(self, scores, nms_window, nfeatures, threshold=0.0):
        """NMS over a float score map (reference NMS(), include/NMS.cuh:5): first `nfeatures` survivors in raster order."""
        sc = np.asarray(scores)
        if sc.dtype != np.float32 or sc.ndim != 2 or sc.strides[1] != 4:
            raise ValueError("scores must be a 2-D float32 array with contiguous rows")
        kps = np.zeros(max(nfeatures, 1), KP)
        n = C.c_int()
        self._ck(self.lib.orb_nms_scores(self.h, _p(sc), sc.shape[1], sc.shape[0], C.c_size_t(sc.strides[0]), int(nms_window), int(nfeatures),
                                         C.c_float(threshold), _p(kps), C.byref(n)))
        return kps[:n.value]

    def conv2d_u8(self, image, kernel, reflect=False, divisor=0.0):
        """conv2d() of the reference (include/Convolution.cuh:5) -> uint8 image; reflect: BORDER_REFLECT_101 first (same size out)."""
        img = _img(image)
        k = np.ascontiguousarray(kernel, np.float32).ravel()
        K = int(round(len(k) ** 0.5))
        h, w = img.shape
        oh, ow = (h, w) if reflect else (h - K + 1, w - K + 1)
        out = np.zeros((max(oh, 1), max(ow, 1)), np.uint8)
        self._ck(self.lib.orb_conv2d_u8(self.h, _p(img), w, h, C.c_size_t(img.strides[0]), _p(k), K, int(bool(reflect)), C.c_float(divisor),
                                        _p(out), C.c_size_t(out.strides[0])))
        return out

    def gaussian_blur_1d(self, image):
        img = _img(image)
        out = np.zeros(img.shape, np.uint8)
        self._ck(self.lib.orb_gaussian_blur_1d(self.h, _p(img), img.shape[1], img.shape[0], C.c_size_t(img.strides[0]), _p(out),
                                               C.c_size_t(out.strides[0])))
        return out

    def harris(self, image, kps):
        img = _img(image)
        kps = np.ascontiguousarray(kps, KP)
        out = np.zeros(len(kps), np.float32)
        self._ck(self.lib.orb_harris(self.h, _p(img), img.shape[1], img.shape[0], img.strides[0], _p(kps), len(kps), _p(out)))
        return out

    def orientations(self, image, kps):
        img = _img(image)
        kps = np.ascontiguousarray(kps, KP)
        out = np.zeros(len(kps), np.float32)
        self._ck(self.lib.orb_orientations(self.h, _p(img), img.shape[1], img.shape[0], img.strides[0], _p(kps),
                                           len(kps), _p(out)))
        return out

    def brief(self, image, kps, angles):
        img = _img(image)
        kps = np.ascontiguousarray(kps, KP)
        angles = np.ascontiguousarray(angles, np.float32)
        if len(angles) != len(kps):
            raise ValueError("keypoints / orientations length mismatch")
        out = np.zeros((len(kps), 32), np.uint8)
        self._ck(self.lib.orb_brief(self.h, _p(img), img.shape[1], img.shape[0], img.strides[0], _p(kps), _p(angles),
                                    len(kps), _p(out)))
        return out

    # ---- descriptor matching (replaces flann->knnMatch(des1, des2, matches, 2) + ratio test of the reference VO loops)
    def match_knn2(self, query, train):
        q = np.ascontiguousarray(query, np.uint8).reshape(-1, 32)
        t = np.ascontiguousarray(train, np.uint8).reshape(-1, 32)
        out = np.zeros(len(q), MATCH)
        self._ck(self.lib.orb_match_knn2(self.h, _p(q), len(q), _p(t), len(t), 0, _p(out)))
        return out

    def match_knn2_batch(self, desc, n):
        desc = np.ascontiguousarray(desc, np.uint8)
        F, cap = desc.shape[0], desc.shape[1]
        n = np.ascontiguousarray(n, np.int32)
        out = np.zeros((max(F - 1, 0), cap), MATCH)
        out["idx1"] = out["idx2"] = -1
        self._ck(self.lib.orb_match_knn2_batch(self.h, _p(desc), _p(n), F, cap, 0, _p(out)))
        return out

    def match_knn2_batch_ptr(self, desc_ptr, n_ptr, n_frames, cap, out_ptr):
        """device-resident form: descriptors / counts as produced by detect_and_compute_batch_ptr(..., outputs_on_device=1)"""
        self._ck(self.lib.orb_match_knn2_batch(self.h, C.c_void_p(desc_ptr), C.c_void_p(n_ptr), n_frames, cap, 1,
                                               C.c_void_p(out_ptr)))

    def ratio_test(self, matches, ratio=0.8):
        m = np.ascontiguousarray(matches, MATCH).ravel()
        keep = np.zeros(len(m), np.uint8)
        self.lib.orb_ratio_test(_p(m), len(m), ratio, _p(keep))
        return keep.astype(bool).reshape(np.shape(matches))

    def bounds_check(self):
        """(enabled, failures, first_line, ctas_checked) of a -DORB_BOUNDS_CHECK build (enabled = 0 in a normal build)."""
        en, fl, ln, kc = C.c_int(), C.c_uint(), C.c_uint(), C.c_uint()
        self._ck(self.lib.orb_debug_bounds_check(self.h, C.byref(en), C.byref(fl), C.byref(ln), C.byref(kc)))
        return en.value, fl.value, ln.value, kc.value

    def bounds_selftest(self):
        self._ck(self.lib.orb_debug_bounds_selftest(self.h))

    def eval_math(self, op, a, b=None):
        a = np.ascontiguousarray(a, np.float32)
        b = None if b is None else np.ascontiguousarray(b, np.float32)
        out = np.zeros(len(a), np.float32)
        self._ck(self.lib.orb_debug_eval_math(self.h, op, _p(a), _p(b), len(a), _p(out)))
        return out


def make_params(nfeatures=500, scaleFactor=1.2, nlevels=8, threshold=20, n=9, nms_window=3, patch_size=31,
                select_policy=SELECT_HARRIS_TOP_N, blur_levels=1, harris_k=0.04, device=0, max_width=1241,
                max_height=376, max_batch=1, chunk_frames=0, max_keypoints=0, keep_side_arrays=0):
    p = default_params()
    p.nfeatures, p.scale_factor, p.nlevels = nfeatures, scaleFactor, nlevels
    p.fast_threshold, p.fast_n, p.nms_window, p.orient_patch = threshold, n, nms_window, patch_size
    p.select_policy, p.blur_levels, p.harris_k = select_policy, blur_levels, harris_k
    p.device, p.max_width, p.max_height, p.max_batch = device, max_width, max_height, max_batch
    p.chunk_frames, p.max_keypoints, p.keep_side_arrays = chunk_frames, max_keypoints, keep_side_arrays
    return p


# ---- the reference's class surface ------------------------------------------------------------
class OrientedFAST:
    """ref include/orb.hpp:10-22: OrientedFAST(threshold=20, n=9, nms_window=3, patch_size=31)."""

    def __init__(self, threshold=20, n=9, nms_window=3, patch_size=31, device=0, max_width=4096, max_height=2304):
        self._ctx = Context(make_params(nfeatures=3000, nlevels=1, threshold=threshold, n=n, nms_window=nms_window,
                                        patch_size=patch_size, select_policy=SELECT_RASTER_FIRST_N, device=device,
                                        max_width=max_width, max_height=max_height))

    def detect(self, image, nfeatures):
        """std::vector<Keypoint> detect(const cv::Mat&, int nfeatures) -- ref src/orb.cpp:22-27."""
        return self._ctx.fast_detect(image, nfeatures)

    def compute_orientations(self, image, keypoints):
        """ref src/orb.cpp:29-33 / src/orb_cpu.cpp:139-183."""
        return self._ctx.orientations(image, keypoints)


class RotatedBRIEF:
    """ref include/orb.hpp:24-32: n_bits = 256, patch_size = 31."""

    def __init__(self, device=0, max_width=4096, max_height=2304):
        self._ctx = Context(make_params(nfeatures=3000, nlevels=1, select_policy=SELECT_RASTER_FIRST_N, device=device,
                                        max_width=max_width, max_height=max_height))

    def compute(self, image, keypoints, orientations):
        """ref src/orb.cpp:40-44 / src/orb_cpu.cpp:203-258."""
        return self._ctx.brief(image, keypoints, orientations)


class ORB:
    """ref include/orb.hpp:34-49: ORB(nfeatures=500, scaleFactor=1.2f, nlevels=8).

    Extra keyword knobs (FAST threshold, patch, selection policy, blur, device, capacities) default to the
    reference's include/orb.hpp values, so ORB() behaves like the reference constructor.
    """

    def __init__(self, nfeatures=500, scaleFactor=1.2, nlevels=8, **kw):
        kw.setdefault("keep_side_arrays", 0)
        self.params = make_params(nfeatures=nfeatures, scaleFactor=scaleFactor, nlevels=nlevels, **kw)
        self.ctx = Context(self.params)

    def detectAndCompute(self, image):
        """void detectAndCompute(image, keypoints, orientations, descriptors) -- ref src/orb.cpp:58-109.
        Returns (keypoints, orientations, descriptors)."""
        k, a, d, _ = self.ctx.detect_and_compute(image)
        return k, a, d

    def detectAndComputeBatch(self, frames, cap=None):
        return self.ctx.detect_and_compute_batch(frames, cap)


class ORBCPU(ORB):
    """Name-compatible stand-in for the reference's CPU twin (include/orb_cpu.hpp:28-43): the SAME GPU
    path configured the way ORBCPU::detectAndCompute actually runs as shipped (single level, FAST thr 50,
    raster-first 3000, patch 9).  It is not a CPU implementation."""

    def __init__(self, nfeatures=500, scaleFactor=1.2, nlevels=8, **kw):
        kw.setdefault("max_keypoints", 3000)
        super().__init__(nfeatures=3000, scaleFactor=scaleFactor, nlevels=1, threshold=50, patch_size=9,
                         select_policy=SELECT_RASTER_FIRST_N, **kw)
