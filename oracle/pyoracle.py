"""ctypes binding of the CPU parity oracle -- TEST INFRASTRUCTURE ONLY.

May be imported by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs, never by the product package.  Wraps oracle/liborb_oracle.so (our restatement,
oracle/orb_oracle.cpp) and, when present, oracle/_ref/liborbcpu_ref.so (the reference's own
src/orb_cpu.cpp compiled unmodified against oracle/cv_shim).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "liborb_oracle.so")
REF_PATH = os.path.join(HERE, "_ref", "liborbcpu_ref.so")

KP = np.dtype([("x", "<i4"), ("y", "<i4")])


class Params(C.Structure):
    _fields_ = [("nfeatures", C.c_int32), ("scale_factor", C.c_float), ("nlevels", C.c_int32),
                ("fast_threshold", C.c_int32), ("fast_n", C.c_int32), ("nms_window", C.c_int32),
                ("orient_patch", C.c_int32), ("select_policy", C.c_int32), ("blur_levels", C.c_int32),
                ("harris_k", C.c_float)]


def params(nfeatures=500, scale_factor=1.2, nlevels=8, fast_threshold=20, fast_n=9, nms_window=3,
           orient_patch=31, select_policy=1, blur_levels=1, harris_k=0.04):
    return Params(nfeatures, scale_factor, nlevels, fast_threshold, fast_n, nms_window, orient_patch,
                  select_policy, blur_levels, harris_k)


def build(force=False):
    """Compile the oracle (and oracle/_ref when /root/reference is present)."""
    if force or not os.path.exists(LIB_PATH) or \
            os.path.getmtime(LIB_PATH) < os.path.getmtime(os.path.join(HERE, "orb_oracle.cpp")):
        subprocess.check_call(["make", "-s", "-C", HERE, "all"])
    elif os.path.exists("/root/reference/src/orb_cpu.cpp") and not os.path.exists(REF_PATH):
        subprocess.check_call(["make", "-s", "-C", HERE, "ref"])


_lib = None
_ref = None


def _u8p(a):
    return a.ctypes.data_as(C.POINTER(C.c_uint8))


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.orc_level_scale.restype = C.c_float
        L.orc_level_scale.argtypes = [C.c_float, C.c_int]
        L.orc_level_quota.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int]
        L.orc_level_size.argtypes = [C.c_int, C.c_int, C.c_float, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_lround_f.restype = C.c_long
        L.orc_lround_f.argtypes = [C.c_float]
        for f in (L.orc_cosf, L.orc_sinf):
            f.restype = C.c_float
            f.argtypes = [C.c_float]
        L.orc_atan2f.restype = C.c_float
        L.orc_atan2f.argtypes = [C.c_float, C.c_float]
        L.orc_pattern.restype = C.POINTER(C.c_int8)
        L.orc_harris.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_int, C.c_float, C.c_void_p]
        L.orc_detect_and_compute.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.POINTER(Params), C.c_int] + \
            [C.c_void_p] * 7
        L.orc_detect_and_compute_batch.argtypes = [C.c_void_p, C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_size_t,
                                                   C.POINTER(Params), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                                   C.c_void_p, C.c_int]
        L.orc_resize_linear_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_int, C.c_int, C.c_size_t]
        L.orc_gauss5x5_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_size_t]
        L.orc_integral_flat.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p]
        L.orc_resize_table.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_build_level.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.POINTER(Params), C.c_int, C.c_void_p]
        L.orc_fast_scores.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_int, C.c_void_p]
        L.orc_nms.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.orc_harris_weights.argtypes = [C.c_void_p]
        L.orc_gaussian_kernel.argtypes = [C.c_int, C.c_void_p]
        L.orc_conv2d_u8.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p]
        L.orc_gaussian_blur_1d.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p]
        L.orc_orientations.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.orc_brief.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.orc_brief_flags.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.orc_match_knn2.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_void_p]
        L.orc_lk_track.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                   C.c_int, C.c_double, C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_lk_levels.argtypes = [C.c_int] * 4
        L.orc_lk_pyr_down.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p]
        _lib = L
    return _lib


def have_ref():
    build()
    return os.path.exists(REF_PATH)


def ref():
    global _ref
    if _ref is None:
        build()
        R = C.CDLL(REF_PATH)
        R.ref_orbcpu_detect_and_compute.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int,
                                                    C.c_void_p, C.c_void_p, C.c_void_p]
        R.ref_fast_detect.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t] + [C.c_int] * 5 + [C.c_void_p]
        R.ref_orientations.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        R.ref_brief.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        R.ref_pattern.restype = C.POINTER(C.c_int)
        _ref = R
    return _ref


def _img(a):
    a = np.ascontiguousarray(a, dtype=np.uint8)
    assert a.ndim == 2
    return a


# ---------------------------------------------------------------- geometry
def level_size(W, H, f, level):
    w, h = C.c_int(), C.c_int()
    lib().orc_level_size(W, H, f, level, C.byref(w), C.byref(h))
    return w.value, h.value


def level_scale(f, level):
    return lib().orc_level_scale(f, level)


def level_quota(nfeatures, f, nlevels, level):
    return lib().orc_level_quota(nfeatures, f, nlevels, level)


def pattern():
    return np.ctypeslib.as_array(lib().orc_pattern(), shape=(1024,)).copy()


# ---------------------------------------------------------------- primitives
def resize_linear(img, dw, dh):
    img = _img(img)
    out = np.empty((dh, dw), np.uint8)
    lib().orc_resize_linear_u8(_ptr(img), img.shape[1], img.shape[0], img.strides[0], _ptr(out), dw, dh, dw)
    return out


def gauss5x5(img):
    img = _img(img)
    out = np.empty_like(img)
    lib().orc_gauss5x5_u8(_ptr(img), img.shape[1], img.shape[0], img.strides[0], _ptr(out), img.shape[1])
    return out


def integral_flat(img):
    img = _img(img)
    h, w = img.shape
    out = np.empty((h + 4, w + 1), np.int32)
    lib().orc_integral_flat(_ptr(img), w, h, img.strides[0], _ptr(out))
    return out


def resize_table(src, dst):
    ofs = np.empty(dst, np.int32)
    a0 = np.empty(dst, np.int16)
    a1 = np.empty(dst, np.int16)
    lib().orc_resize_table(src, dst, _ptr(ofs), _ptr(a0), _ptr(a1))
    return ofs, a0, a1


def build_level(img, p, level):
    img = _img(img)
    H, W = img.shape
    w, h = level_size(W, H, p.scale_factor, level)
    out = np.empty((h, w), np.uint8)
    lib().orc_build_level(_ptr(img), W, H, img.strides[0], C.byref(p), level, _ptr(out))
    return out


# ---------------------------------------------------------------- stages
def fast_scores(img, thr, n=9):
    img = _img(img)
    h, w = img.shape
    out = np.empty((h, w), np.float32)
    lib().orc_fast_scores(_ptr(img), w, h, img.strides[0], thr, n, _ptr(out))
    return out


def nms(scores, nms_window=3, cap=None):
    scores = np.ascontiguousarray(scores, np.float32)
    h, w = scores.shape
    cap = w * h if cap is None else cap
    kps = np.empty(max(cap, 1), KP)
    n = lib().orc_nms(_ptr(scores), w, h, nms_window, cap, _ptr(kps))
    return kps[:n].copy()


def harris_weights():
    w = np.empty(49, np.float32)
    lib().orc_harris_weights(_ptr(w))
    return w


def gaussian_kernel(ksize):
    """createGaussianKernel(ksize) of the reference (src/GaussianBlur.cpp:7-37)."""
    k = np.empty(ksize * ksize, np.float32)
    lib().orc_gaussian_kernel(ksize, _ptr(k))
    return k


def conv2d_u8(img, kernel, reflect=False, divisor=0.0):
    """conv2d() of the reference (src/cuda/Convolution.cu): valid-mode correlation -> CV_8U; reflect: BORDER_REFLECT_101 first."""
    img = _img(img)
    kernel = np.ascontiguousarray(kernel, np.float32).ravel()
    K = int(round(len(kernel) ** 0.5))
    h, w = img.shape
    oh, ow = (h, w) if reflect else (h - K + 1, w - K + 1)
    out = np.empty((oh, ow), np.uint8)
    lib().orc_conv2d_u8(_ptr(img), w, h, img.strides[0], _ptr(kernel), K, int(reflect), float(divisor), _ptr(out))
    return out


def gaussian_blur_1d(img):
    img = _img(img)
    out = np.empty(img.shape, np.uint8)
    lib().orc_gaussian_blur_1d(_ptr(img), img.shape[1], img.shape[0], img.strides[0], _ptr(out))
    return out


def harris(img, kps, k=0.04):
    img = _img(img)
    kps = np.ascontiguousarray(kps, KP)
    out = np.empty(len(kps), np.float32)
    lib().orc_harris(_ptr(img), img.shape[1], img.shape[0], img.strides[0], _ptr(kps), len(kps), k, _ptr(out))
    return out


def orientations(img, kps, patch):
    img = _img(img)
    kps = np.ascontiguousarray(kps, KP)
    out = np.empty(len(kps), np.float32)
    lib().orc_orientations(_ptr(img), img.shape[1], img.shape[0], img.strides[0], _ptr(kps), len(kps), patch, _ptr(out))
    return out


def brief(img, kps, angles):
    img = _img(img)
    kps = np.ascontiguousarray(kps, KP)
    angles = np.ascontiguousarray(angles, np.float32)
    out = np.zeros((len(kps), 32), np.uint8)
    lib().orc_brief(_ptr(img), img.shape[1], img.shape[0], img.strides[0], _ptr(kps), _ptr(angles), len(kps), _ptr(out))
    return out


def brief_flags(w, h, kps, angles):
    kps = np.ascontiguousarray(kps, KP)
    angles = np.ascontiguousarray(angles, np.float32)
    out = np.zeros(len(kps), np.uint8)
    lib().orc_brief_flags(w, h, _ptr(kps), _ptr(angles), len(kps), _ptr(out))
    return out


def match_knn2(query, train, ratio=0.8):
    """Exact Hamming 2-NN per query: (int32[nq,4] = idx1, dist1, idx2, dist2; bool[nq] ratio test)."""
    q = np.ascontiguousarray(query, np.uint8).reshape(-1, 32)
    t = np.ascontiguousarray(train, np.uint8).reshape(-1, 32)
    out = np.zeros((len(q), 4), np.int32)
    keep = np.zeros(len(q), np.uint8)
    lib().orc_match_knn2(_ptr(q), len(q), _ptr(t), len(t), _ptr(out), ratio, _ptr(keep))
    return out, keep.astype(bool)


def lk_track(prev, nxt, pts, win=21, max_level=3, max_iter=30, eps=0.01, min_eig=0.001):
    """== cv2.calcOpticalFlowPyrLK(prev, nxt, pts, None, winSize=(win, win), maxLevel, criteria, 0, min_eig):
    returns (next_pts [n,2] float32, status [n] uint8, err [n] float32)."""
    prev, nxt = _img(prev), _img(nxt)
    assert prev.shape == nxt.shape and prev.strides[0] == nxt.strides[0]
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
    out = np.zeros_like(pts)
    st = np.zeros(len(pts), np.uint8)
    er = np.zeros(len(pts), np.float32)
    lib().orc_lk_track(_ptr(prev), _ptr(nxt), prev.shape[1], prev.shape[0], prev.strides[0], _ptr(pts), len(pts), win, max_level,
                       max_iter, eps, min_eig, _ptr(out), _ptr(st), _ptr(er))
    return out, st, er


def lk_pyr_down(img):
    img = _img(img)
    h, w = img.shape
    out = np.zeros(((h + 1) // 2, (w + 1) // 2), np.uint8)
    lib().orc_lk_pyr_down(_ptr(img), w, h, img.strides[0], _ptr(out))
    return out


# ---------------------------------------------------------------- whole path
def detect_and_compute(img, p, cap=None):
    """Returns dict(kps, angles, desc, n_per_level, level_xy, level_id, response)."""
    img = _img(img)
    H, W = img.shape
    cap = cap if cap is not None else max(1, p.nfeatures * (p.nlevels if p.select_policy == 0 else 1))
    kps = np.zeros(cap, KP)
    ang = np.zeros(cap, np.float32)
    des = np.zeros((cap, 32), np.uint8)
    npl = np.zeros(p.nlevels, np.int32)
    lxy = np.zeros(cap, KP)
    lid = np.zeros(cap, np.int32)
    rsp = np.zeros(cap, np.float32)
    n = lib().orc_detect_and_compute(_ptr(img), W, H, img.strides[0], C.byref(p), cap, _ptr(kps), _ptr(ang),
                                     _ptr(des), _ptr(npl), _ptr(lxy), _ptr(lid), _ptr(rsp))
    return dict(n=n, kps=kps[:n], angles=ang[:n], desc=des[:n], n_per_level=npl, level_xy=lxy[:n],
                level_id=lid[:n], response=rsp[:n])


def detect_and_compute_batch(frames, p, cap, n_threads, keep=False):
    """frames: (F,H,W) or (F,H,pitch) uint8 contiguous; returns n_out (and records if keep)."""
    frames = np.ascontiguousarray(frames, np.uint8)
    F, H, W = frames.shape
    n_out = np.zeros(F, np.int32)
    kps = ang = des = None
    if keep:
        kps = np.zeros((F, cap), KP)
        ang = np.zeros((F, cap), np.float32)
        des = np.zeros((F, cap, 32), np.uint8)
    lib().orc_detect_and_compute_batch(_ptr(frames), F, frames.strides[0], W, H, frames.strides[1], C.byref(p), cap,
                                       _ptr(kps), _ptr(ang), _ptr(des), _ptr(n_out), n_threads)
    return (n_out, kps, ang, des) if keep else n_out


# ---------------------------------------------------------------- the reference itself (oracle/_ref)
def ref_orbcpu(img, cap=3000):
    """ORBCPU().detectAndCompute on img: the shipped single-level path (3000/50/9/3/patch 9)."""
    img = _img(img)
    kps = np.zeros(cap, KP)
    ang = np.zeros(cap, np.float32)
    des = np.zeros((cap, 32), np.uint8)
    n = ref().ref_orbcpu_detect_and_compute(_ptr(img), img.shape[1], img.shape[0], img.strides[0], cap,
                                            _ptr(kps), _ptr(ang), _ptr(des))
    n = min(n, cap)
    return kps[:n], ang[:n], des[:n]


def ref_fast_detect(img, nfeatures, thr, n=9, nms=3):
    img = _img(img)
    cap = max(nfeatures, 1)
    kps = np.zeros(cap, KP)
    m = ref().ref_fast_detect(_ptr(img), img.shape[1], img.shape[0], img.strides[0], nfeatures, thr, n, nms, cap, _ptr(kps))
    return kps[:min(m, cap)].copy()


def ref_orientations(img, kps, patch):
    img = _img(img)
    kps = np.ascontiguousarray(kps, KP)
    out = np.empty(len(kps), np.float32)
    ref().ref_orientations(_ptr(img), img.shape[1], img.shape[0], img.strides[0], _ptr(kps), len(kps), patch, _ptr(out))
    return out


def ref_brief(img, kps, angles):
    img = _img(img)
    kps = np.ascontiguousarray(kps, KP)
    angles = np.ascontiguousarray(angles, np.float32)
    out = np.zeros((len(kps), 32), np.uint8)
    ref().ref_brief(_ptr(img), img.shape[1], img.shape[0], img.strides[0], _ptr(kps), _ptr(angles), len(kps), _ptr(out))
    return out


def ref_pattern():
    return np.ctypeslib.as_array(ref().ref_pattern(), shape=(1024,)).copy()


def fnv1a64(arr):
    """FNV-1a-64 over the raw little-endian bytes of an array (the survey's known-answer hash)."""
    data = np.ascontiguousarray(arr).tobytes()
    h = 0xcbf29ce484222325
    for b in data:
        h = ((h ^ b) * 0x100000001b3) & 0xFFFFFFFFFFFFFFFF
    return "%016x" % h
